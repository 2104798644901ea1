#!/usr/bin/env python
"""train.py (reference, lines 180-281) end to end on the drop-in: the reference's on-disk files -> HigherDataset ->
device graph build -> ablation3 / Ours -> nll_loss -> Adam, then the test pass with the reference's metrics.

    python tools/train_flow.py --root /path/to/anonymous_data --year 2015 --model ablation3 --epochs 1
    python tools/train_flow.py --synthetic /tmp/flow2015        # writes 2015-shaped files first (no reference data on the box)

Only the imports differ from train.py: ``dataset.HigherDataset`` -> ``msha_gnn_b200.HigherDataset`` (reads the same
files, never builds the (N, N) matrices), ``Ablation`` / ``Ours`` -> ``msha_gnn_b200``.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F
import torch.optim as optim
from torch.utils.data import DataLoader, random_split

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import msha_gnn_b200 as mg                      # noqa: E402
from msha_gnn_b200.data import write_flow_files  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--root", default=os.environ.get("MSHA_DATA_ROOT", "/root/reference/anonymous_data"))
    ap.add_argument("--year", default="2015")
    ap.add_argument("--synthetic", default=None, help="directory to write a 2015-shaped synthetic year into and train on")
    ap.add_argument("--model", default="ablation3", choices=["ablation3", "ablation2", "Ours"])
    ap.add_argument("--seed", type=int, default=42)                 # train.py:25
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--lr", type=float, default=0.001)              # train.py:29
    ap.add_argument("--weight_decay", type=float, default=5e-4)     # train.py:31
    ap.add_argument("--batch_size", type=int, default=64)           # train.py:33
    ap.add_argument("--max_steps", type=int, default=0, help="stop an epoch early (0 = full epoch)")
    ap.add_argument("--cuda-graph", action="store_true",
                    help="replay the training step as one CUDA graph (msha_gnn_b200.CapturedStep); full batches only")
    args = ap.parse_args()
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    device = torch.device("cuda")
    if args.synthetic:
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from bench import flow_graph
        src, dst, city, prov = flow_graph()
        write_flow_files(args.synthetic, args.year, src, dst, city, prov, np.random.default_rng(0).random(city.size),
                         recipient_names=[f"r{j}" for j in range(32)])
        args.root = args.synthetic
    t1 = time.time()
    Dataset = mg.HigherDataset(args.root, args.year, device=device)                         # train.py:180
    train_size = int(0.9 * len(Dataset))
    train_dataset, test_dataset = random_split(Dataset, [train_size, len(Dataset) - train_size])
    train_loader = DataLoader(train_dataset, batch_size=args.batch_size, shuffle=True)
    test_loader = DataLoader(test_dataset, batch_size=args.batch_size, shuffle=False)
    Scount, Rcount = Dataset.get_count()
    inter_adj, city_adj, province_adj = (mg.normalize_adjacency_matrix(a) for a in Dataset.get_adjacent())   # train.py:190-193
    GDP = Dataset.get_gdp()
    torch.cuda.synchronize()
    print("load data: {:.3f}s  (N={}, M={}, records={}, nnz={})".format(time.time() - t1, Scount, Rcount, len(Dataset), inter_adj.nnz))
    model = getattr(mg, args.model)(in_features=128, out_features=64, n_classes=Rcount, n_heads=2, dropout=0.5, gdp=GDP,
                                    Scount=Scount, Rcount=Rcount).to(device)                # train.py:206
    optimizer = optim.Adam(model.parameters(), lr=args.lr, weight_decay=args.weight_decay,
                           capturable=args.cuda_graph)                                       # train.py:207

    def train_step(source_index, recipient_index):                                          # train.py:221-232
        optimizer.zero_grad(set_to_none=True)
        output = model(inter_adj, city_adj, province_adj, source_index)
        loss_train = F.nll_loss(output[source_index], recipient_index)                      # train.py:229
        loss_train.backward()
        optimizer.step()
        return loss_train

    captured = None
    for epoch in range(args.epochs):
        t = time.time()
        model.train()
        Loss_train, n = 0.0, 0
        for i, (source_index, recipient_index) in enumerate(train_loader):
            source_index, recipient_index = source_index.to(device), recipient_index.to(device)
            if args.cuda_graph and source_index.numel() == args.batch_size:
                if captured is None:
                    captured = mg.CapturedStep(train_step, [source_index, recipient_index])
                loss_train = captured(source_index, recipient_index)
            else:
                loss_train = train_step(source_index, recipient_index)
            Loss_train += loss_train.item()
            n += 1
            if args.max_steps and n >= args.max_steps:
                break
        print('Epoch: {:04d}'.format(epoch + 1), 'loss_train: {:.4f}'.format(Loss_train / n),
              'time: {:.4f}s'.format(time.time() - t), '({:.2f} ms/step)'.format((time.time() - t) / n * 1e3))
        model.eval()
        correct = total = 0
        Loss_test, m = 0.0, 0
        with torch.no_grad():
            for source_index, recipient_index in test_loader:
                source_index, recipient_index = source_index.to(device), recipient_index.to(device)
                output = model(inter_adj, city_adj, province_adj, source_index)
                Loss_test += F.nll_loss(output[source_index], recipient_index).item()
                pred = output[source_index].max(1)[1]
                correct += int((pred == recipient_index).sum())
                total += recipient_index.numel()
                m += 1
                if args.max_steps and m >= args.max_steps:
                    break
        print("Test set results:", "loss= {:.4f}".format(Loss_test / m), "accuracy= {:.4f}".format(correct / total))


if __name__ == "__main__":
    main()
