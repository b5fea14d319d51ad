"""A small training step (2 Trainer steps, 4 x 64 x 96 images) for compute-sanitizer: every csnet_train_* kernel runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sod100k_b200 import checkpoints, synth
from sod100k_b200.trainer import Trainer

model, cfg, _ = checkpoints.build_from_npz("csnet-L-x2")
model.cuda(0)
tr = Trainer(model, lr=1e-4, weight_decay=5e-3, flops_weight=3.0)
x = torch.from_numpy(synth.randn_images(4, 64, 96, 3)).cuda()
t = torch.from_numpy(synth.random_masks(4, 64, 96, 4)).cuda()
for _ in range(2):
    loss = tr.step(x, t)
torch.cuda.synchronize()
print("loss", float(loss))
