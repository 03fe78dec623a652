#!/usr/bin/env python3
"""Zero-cost switching: every step uses another (adaptive_r, a2a_ffn_overlap_degree) pair on the SAME parameters
(reference: tutel/examples/helloworld_switch.py).   python -m tutel_b200.examples.helloworld_switch --batch_size=16"""
from tutel_b200.examples._driver import MoEClassifier, Session, base_parser, default_layer, manual_allreduce


def main(argv=None):
    args = base_parser().parse_args(argv)
    s = Session(args)
    layer = default_layer(s, gate_type={'type': 'top', 'k': args.top, 'fp32_gate': args.fp32_gate})
    s.report_params(layer)
    state = {'i': -1}

    def call(layer, x):
        rs = layer.valid_rs
        r, o = rs[(state['i'] // 8) % len(rs)], state['i'] % 8 + 1
        state['i'] += 1
        return layer(x, capacity_factor=args.cap_factor, adaptive_r=r, a2a_ffn_overlap_degree=o)

    model = MoEClassifier(layer, call).to(s.device)
    s.print(model)
    import torch
    opt = torch.optim.SGD(model.parameters(), lr=1e-5)
    x, y = s.synthetic_batch()
    s.banner()
    s.train(model, opt, x, y, sync_grads=manual_allreduce(s, model),
            suffix=lambda m: '(f = %.1f, r = %d, o = %d)' % (args.cap_factor, m._moe_layer.adaptive_degree, m._moe_layer.a2a_ffn_overlap_degree))


if __name__ == '__main__':
    main()
