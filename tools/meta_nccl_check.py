"""torchrun --nproc-per-node N tools/meta_nccl_check.py : one meta-iteration (config 5) with N ranks, one task per
rank, ONE NCCL all-reduce of the flat meta-gradient; checks every rank ends with bit-identical MetaOptimizer
parameters (the replicated fused RAdam needs no broadcast)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import numpy as np
import bench
from eosvos_b200.util import meta_train, synthetic

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
model, opt = bench.build_model(dev)
model.transform.min_size, model.transform.max_size = (320,), 533
radam = meta_train.FusedRAdam(opt)
frames, labels = synthetic.make_video(100 + rank, 2, 192, 342, 1)          # every rank: its own task
fr = torch.from_numpy(frames).permute(0, 3, 1, 2).float().div(255.0).to(dev)
gts = torch.from_numpy((labels == 1).astype(np.float32))[:, None].to(dev)
task = ((fr[0:1], gts[0:1]), (fr[1:2], gts[1:2]))
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
losses = meta_train.meta_iteration(model, opt, radam, [task], meta_batch_size=world, num_epochs=5, bptt_epochs=5)
torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
chk = torch.stack([p.detach().double().sum() for _, p in opt.named_parameters()]).sum().reshape(1)
absum = torch.stack([p.detach().double().abs().sum() for _, p in opt.named_parameters()]).sum().reshape(1)
both = torch.cat([chk, absum])
gathered = [torch.zeros_like(both) for _ in range(world)]
dist.all_gather(gathered, both)
if rank == 0:
    same = all(torch.equal(g, gathered[0]) for g in gathered)
    print(f"meta-iteration with {world} rank(s): {dt*1e3:.0f} ms; train/meta losses rank0 {[float(x) for x in losses[0]]}; "
          f"parameters identical across ranks: {same}; checksum {gathered[0].tolist()}")
    assert same
dist.destroy_process_group()
