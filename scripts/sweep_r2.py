"""Round-2 sweep of the tile-plan knobs (CTA budget, shortest tile, mixed-column cost) at the small meshes;
one process, CUDA-event timings (plain launches).  usage: python scripts/sweep_r2.py [ldg|ip] ORDER..."""
import os
import sys

import torch

sys.path.insert(0, ".")
from scripts.probe_small import ip_slab, ldg_op, time_fn  # noqa: E402


def run(kind, order, env):
    for k, v in env.items():
        os.environ[k] = str(v)
    if kind == "ldg":
        op, x = ldg_op(order)
        y = torch.empty_like(x)
        fn = lambda: op.apply(x, out=y)  # noqa: E731
    else:
        op, x, xlo, xhi = ip_slab(order, 8)
        y = torch.empty_like(x)
        fn = lambda: op.apply(x, out=y, x_lo=xlo, x_hi=xhi)  # noqa: E731
    ms = time_fn(fn, 100)
    for k in env:
        del os.environ[k]
    return op.ndofs / ms / 1e6, ms * 1e3


if __name__ == "__main__":
    kind = sys.argv[1]
    for order in [int(a) for a in sys.argv[2:]]:
        for budget in (296,):
            row = []
            for mq in (0, 1, 4, 8, 16):
                g, us = run(kind, order, {"STEFAN_CTA_BUDGET": budget, "STEFAN_MIXED_COST_16": mq})
                row.append(f"mq{mq}: {us:6.2f} us {g:6.1f} G")
            print(f"{kind} p={order} budget {budget}: " + "   ".join(row), flush=True)
