"""Development aid: the roofline_builders block of bench.py on its own."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from hybdrt_b200 import engine as E  # noqa: E402
from hybdrt_b200.models import DRT  # noqa: E402

eng = E.get_engine(0)
drt = DRT()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
print(json.dumps(bench.builder_rooflines(eng, drt, flush, 6553.0), indent=1))
