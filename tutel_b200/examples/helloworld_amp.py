#!/usr/bin/env python3
"""Mixed precision: fp32 parameters, autocast compute, GradScaler (reference: tutel/examples/helloworld_amp.py)."""
import torch

from tutel_b200.examples._driver import MoEClassifier, Session, base_parser, default_layer, manual_allreduce


def main(argv=None):
    p = base_parser()
    p.add_argument('--amp_dtype', type=str, default='float16')
    args = p.parse_args(argv)
    s = Session(args)
    layer = default_layer(s, gate_type={'type': 'top', 'k': args.top, 'fp32_gate': args.fp32_gate})
    s.report_params(layer)
    model = MoEClassifier(layer).to(s.device)
    s.print(model)
    opt = torch.optim.SGD(model.parameters(), lr=1e-5)
    x, y = s.synthetic_batch()
    s.banner()
    amp_dtype = torch.float16 if args.amp_dtype == 'float16' else torch.bfloat16
    dev_type = s.device.type
    scaler = torch.amp.GradScaler(dev_type, enabled=(dev_type == 'cuda' and amp_dtype == torch.float16))

    def forward(m, inp):
        with torch.amp.autocast(dev_type, dtype=amp_dtype if dev_type == 'cuda' else torch.bfloat16):
            return m(inp).float()

    s.train(model, opt, x, y, sync_grads=manual_allreduce(s, model), forward=forward, scaler=scaler)


if __name__ == '__main__':
    main()
