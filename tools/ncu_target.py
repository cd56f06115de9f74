"""Short, deterministic workload for ncu: 1 warm-up step + merge + N steps of llama_350m SoW r=50 at batch B."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sow_b200.trainer import SoWTrainer, TrainConfig  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
tr = SoWTrainer(TrainConfig(batch_size=B), dev)
ids = torch.randint(1, 32000, (B, 256), device=dev)
tr.step(ids)
tr.merge()
for _ in range(N):
    loss = tr.step(ids)
tr.merge()
torch.cuda.synchronize()
print("done", float(loss))
