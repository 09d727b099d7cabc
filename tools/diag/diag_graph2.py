import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tests.test_graph_gpu import _run
from tools.diag.diag_graph import maxdiff  # noqa  (runs its own comparison on import; cheap enough)
