"""One full-size training step (B = 240, 12 + 6 + 6 layers) and nothing else: the process `ncu --metrics gpu__time_duration.sum` lists
the kernels of (profiles/r02_v6_train_launches.txt)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unimm_b200 import synthetic as syn  # noqa: E402
from unimm_b200.config import DEFAULT_CONFIG_PATH, ViLBertConfig  # noqa: E402
from unimm_b200.train_ops import DeviceOps  # noqa: E402
from unimm_b200.train_step import TrainStep  # noqa: E402
from unimm_b200.weights import random_state_dict  # noqa: E402

cfg = ViLBertConfig.from_json_file(DEFAULT_CONFIG_PATH)
ts = TrainStep(cfg, random_state_dict(cfg, 0), DeviceOps("cuda:0", sys.argv[1] if len(sys.argv) > 1 else "fp16"))
b = dict(syn.train_batch(1000), nsp_weight=np.array([5.0, 1.0], np.float32))
print(ts.step(b))
torch.cuda.synchronize()
