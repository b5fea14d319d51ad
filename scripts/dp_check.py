"""torchrun --nproc-per-node 2 scripts/dp_check.py — data-parallel gradient check on 2 GPUs (NCCL).
Each rank back-propagates its shard of a global batch with BatchNorm frozen (eval statistics, so BN locality cannot
confound); the flat gradient bucket is all-reduced (mean); rank 0 compares with the single-process gradient of the whole
batch.  Then one Trainer step under DP must leave both ranks with bit-identical parameters."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from sod100k_b200 import checkpoints, synth, train_ops as T
from sod100k_b200.trainer import FlatGrads, Trainer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
per = 2
x = torch.from_numpy(synth.randn_images(per * world, 64, 64, 5)).cuda()
t = torch.from_numpy(synth.random_masks(per * world, 64, 64, 6)).cuda()

def grads_of(xb, tb):
    m, _, _ = checkpoints.build_from_npz("csnet-L-x2")
    m.cuda().eval()
    m.frozen_bn_training = True
    flat = FlatGrads(m.parameters())
    loss = T.BceFn.apply(m(xb), tb)
    loss.backward()
    assert flat.intact()
    return flat

flat = grads_of(x[rank * per:(rank + 1) * per], t[rank * per:(rank + 1) * per])
flat.all_reduce_mean()
ok = True
if rank == 0:
    ref = grads_of(x, t)
    scale = ref.bucket.abs().max().item()
    err = (flat.bucket - ref.bucket).abs().max().item() / scale
    print(f"dp_check: world={world} bucket={ref.bucket.numel()} floats ({ref.bucket.numel() * 4} bytes), "
          f"max |allreduced - single-process| / max|g| = {err:.3e}")
    ok = err <= 1e-4
# a full DP training step: parameters must stay identical across ranks
m, _, _ = checkpoints.build_from_npz("csnet-L-x2")
m.cuda()
tr = Trainer(m, lr=1e-4)
tr.step(x[rank * per:(rank + 1) * per], t[rank * per:(rank + 1) * per])
vec = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
others = [torch.empty_like(vec) for _ in range(world)]
dist.all_gather(others, vec)
same = all(torch.equal(o, others[0]) for o in others)
if rank == 0:
    print("dp_check: parameters identical across ranks after one DP step:", same)
    print("dp_check:", "PASS" if (ok and same) else "FAIL")
dist.destroy_process_group()
sys.exit(0 if (ok and same) else 1)
