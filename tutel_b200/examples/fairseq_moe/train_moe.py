#!/usr/bin/env python3
"""fairseq-train with tutel_b200 MoE layers: wraps ``fairseq_cli.train`` so that models are converted right after they
are built and the auxiliary losses join the training loss - the two hooks the reference implements by patching
fairseq's sources (tutel/examples/fairseq_moe/fairseq_patch.diff:24-128)."""
import sys


def main():
    try:
        from fairseq import tasks
        from fairseq_cli import train as fairseq_train
    except ImportError:
        sys.exit('fairseq is not installed: `pip install fairseq` (or add it to PYTHONPATH) to use this launcher.')
    from tutel_b200.examples.fairseq_moe import add_moe_aux_loss, convert_transformer_layers, zero_overflow_grads

    base = tasks.FairseqTask
    build_model, train_step = base.build_model, base.train_step

    def build_model_moe(self, cfg, *args, **kwargs):
        model = build_model(self, cfg, *args, **kwargs)
        n = convert_transformer_layers(model)
        if n:
            print('[tutel_b200] converted %d transformer FFN(s) to MoE layers' % n, flush=True)
        return model

    def train_step_moe(self, sample, model, criterion, optimizer, update_num, ignore_grad=False):
        # identical to FairseqTask.train_step except for the auxiliary loss and the overflow guard
        import torch
        model.train()
        model.set_num_updates(update_num)
        with torch.autograd.profiler.record_function('forward'):
            loss, sample_size, logging_output = criterion(model, sample)
        loss = add_moe_aux_loss(loss)
        if ignore_grad:
            loss *= 0
        with torch.autograd.profiler.record_function('backward'):
            optimizer.backward(loss)
        zero_overflow_grads(p for g in getattr(optimizer, 'param_groups', []) for p in g['params'])
        return loss, sample_size, logging_output

    base.build_model, base.train_step = build_model_moe, train_step_moe
    fairseq_train.cli_main()


if __name__ == '__main__':
    main()
