"""Shared trainer of the image-classification MoE examples (MNIST / CIFAR-10).

A small conv-net whose classifier head is either a dense 2-layer MLP or a Tutel-style MoE layer (one expert per
device, dropless routing, gate noise).  The datasets are read through torchvision when they are available on disk;
otherwise (no network in many clusters) a deterministic synthetic stand-in with the same shapes is used so that the
example always runs.
"""
import argparse

import torch
import torch.nn as nn
import torch.nn.functional as F

from tutel_b200 import moe, net, system


def load_dataset(name, train, root='/tmp/data'):
    shape = {'mnist': (1, 28, 28), 'cifar10': (3, 32, 32)}[name]
    try:
        from torchvision import datasets, transforms
        cls = datasets.MNIST if name == 'mnist' else datasets.CIFAR10
        norm = transforms.Normalize((0.1307,), (0.3081,)) if name == 'mnist' else transforms.Normalize((0.5,) * 3, (0.5,) * 3)
        return cls(root, train=train, download=False, transform=transforms.Compose([transforms.ToTensor(), norm]))
    except Exception:  # noqa
        g = torch.Generator().manual_seed(0 if train else 1)
        n = 2048 if train else 512
        labels = torch.randint(0, 10, (n,), generator=g)
        protos = torch.randn(10, *shape, generator=torch.Generator().manual_seed(7))
        images = protos[labels] + 0.5 * torch.randn(n, *shape, generator=g)
        return torch.utils.data.TensorDataset(images, labels)


class Net(nn.Module):
    def __init__(self, dataset, use_moe, env):
        super().__init__()
        in_ch, feat = (1, 9216) if dataset == 'mnist' else (3, 12544)
        self.use_moe = use_moe
        self.dropout1, self.dropout2 = nn.Dropout(0.25), nn.Dropout(0.5)
        if use_moe:
            self.moe_ffn = moe.moe_layer(
                gate_type={'type': 'top', 'k': 1, 'capacity_factor': 0, 'gate_noise': 1.0},
                experts={'type': 'ffn', 'num_experts_per_device': 1, 'hidden_size_per_expert': 128, 'output_dim': 10,
                         'activation_fn': lambda x: self.dropout2(F.relu(x))},
                model_dim=feat, seeds=(1, env.global_rank + 1),
                scan_expert_func=lambda name, param: setattr(param, 'skip_allreduce', True))
        else:
            torch.manual_seed(1)
            self.fc1, self.fc2 = nn.Linear(feat, 128), nn.Linear(128, 10)
        torch.manual_seed(1)
        self.conv1, self.conv2 = nn.Conv2d(in_ch, 32, 3, 1), nn.Conv2d(32, 64, 3, 1)

    def forward(self, x, top_k=None):
        x = F.relu(self.conv1(x))
        x = F.max_pool2d(F.relu(self.conv2(x)), 2)
        x = torch.flatten(self.dropout1(x), 1)
        if self.use_moe:
            x = self.moe_ffn(x, top_k=top_k)
        else:
            x = self.fc2(self.dropout2(F.relu(self.fc1(x))))
        return F.log_softmax(x, dim=1)


def run(dataset, argv=None):
    ap = argparse.ArgumentParser(description='%s example with an MoE classifier head' % dataset.upper())
    ap.add_argument('--batch-size', type=int, default=64)
    ap.add_argument('--test-batch-size', type=int, default=1000)
    ap.add_argument('--epochs', type=int, default=20)
    ap.add_argument('--lr', type=float, default=1.0)
    ap.add_argument('--gamma', type=float, default=0.7)
    ap.add_argument('--dry-run', action='store_true', default=False)
    ap.add_argument('--log-interval', type=int, default=10)
    ap.add_argument('--no-moe', action='store_true', default=False)
    ap.add_argument('--save-model', action='store_true', default=False)
    args = ap.parse_args(argv)

    env = system.init_data_model_parallel(backend='nccl' if torch.cuda.is_available() else 'gloo')
    device = env.local_device
    torch.manual_seed(1)
    train_set, test_set = load_dataset(dataset, True), load_dataset(dataset, False)
    sampler = torch.utils.data.distributed.DistributedSampler(train_set, num_replicas=env.global_size, rank=env.global_rank) \
        if env.global_size > 1 else None
    train_loader = torch.utils.data.DataLoader(train_set, batch_size=args.batch_size, sampler=sampler, shuffle=sampler is None, drop_last=True)
    test_loader = torch.utils.data.DataLoader(test_set, batch_size=args.test_batch_size)

    model = Net(dataset, not args.no_moe, env).to(device)
    optimizer = torch.optim.Adadelta(model.parameters(), lr=args.lr)
    scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=1, gamma=args.gamma)
    shared = [p for p in model.parameters() if not hasattr(p, 'skip_allreduce')]

    for epoch in range(1, args.epochs + 1):
        model.train()
        for i, (data, target) in enumerate(train_loader):
            data, target = data.to(device), target.to(device)
            optimizer.zero_grad()
            loss = F.nll_loss(model(data), target)
            if not args.no_moe:
                loss = loss + 0.0001 * model.moe_ffn.l_aux
            loss.backward()
            if env.global_size > 1:
                for p in shared:
                    p.grad = net.simple_all_reduce(p.grad) / env.global_size
            optimizer.step()
            if i % args.log_interval == 0:
                env.dist_print('Train Epoch: %d [%d/%d]\tLoss: %.6f' % (epoch, i * len(data), len(train_loader.dataset), loss.item()))
            if args.dry_run:
                break
        model.eval()
        correct = total = 0
        with torch.no_grad():
            for data, target in test_loader:
                data, target = data.to(device), target.to(device)
                if env.global_size > 1 and not args.no_moe and data.size(0) % 1:
                    continue
                pred = model(data).argmax(dim=1)
                correct += int((pred == target).sum())
                total += int(target.numel())
        env.dist_print('\nTest set (rank 0 view): Accuracy: %d/%d (%.1f%%)\n' % (correct, total, 100.0 * correct / max(total, 1)))
        scheduler.step()
        if args.dry_run:
            break
    if args.save_model:
        torch.save(model.state_dict(), system.apply_rank_size_from_pattern('%s_moe_{rank}-of-{size}.pt' % dataset, env.global_rank, env.global_size))
