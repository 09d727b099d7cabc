"""Diagnostic: max parameter difference eager vs graph (and eager vs eager) after 7 steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tests.test_graph_gpu import _run


def maxdiff(a, b):
    worst = (0.0, None)
    for (k, p), (_, q) in zip(a.gen.named_parameters(), b.gen.named_parameters()):
        d = float((p - q).abs().max())
        if d > worst[0]:
            worst = (d, k)
    return worst


se, le = _run(False, 7, "bf16")
se2, le2 = _run(False, 7, "bf16")
sg, lg = _run(True, 7, "bf16")
print(os.environ.get("TAG", ""), "eager-eager", maxdiff(se, se2), "eager-graph", maxdiff(se, sg), flush=True)
print(" losses e", le[-1], "\n losses g", lg[-1])
