"""Run under torchrun with >= 2 GPUs:  data-parallel training step == single-GPU step on the concatenated batch.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multigpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cgvc  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    per = 2
    prec = sys.argv[1] if len(sys.argv) > 1 else "f16f8"
    rs = np.random.RandomState(0)
    A = rs.randn(per * world, 24, 128).astype(np.float32); B = rs.randn(per * world, 24, 128).astype(np.float32)
    m = cgvc.CycleGAN(24, mode='train', max_batch=per, max_frames=128, precision=prec, device=lr, seed=123, data_parallel=True, log_dir='/tmp/cgvc_log')
    g, d = m.train(A[rank * per:(rank + 1) * per], B[rank * per:(rank + 1) * per], 10.0, 5.0, 2e-4, 1e-4)
    losses = torch.tensor([float(g), float(d)], device="cuda")
    dist.all_reduce(losses); losses /= world
    after = m.get_params()
    if rank == 0:
        ref = cgvc.CycleGAN(24, mode='train', max_batch=per * world, max_frames=128, precision=prec, device=lr, seed=123, log_dir='/tmp/cgvc_log')
        before = ref.get_params()
        g1, d1 = ref.train(A, B, 10.0, 5.0, 2e-4, 1e-4)
        single = ref.get_params()
        worst = 0.0
        for k in single:
            if "bias" in k and "block" in k:
                continue
            du = (single[k] - before[k]).astype(np.float64); dd = (after[k] - before[k]).astype(np.float64)
            e = np.linalg.norm(du - dd) / (np.linalg.norm(du) + 1e-30)
            worst = max(worst, e)
        print("MULTIGPU world=%d [%s]: mean loss G %.6f (single %.6f) D %.6f (single %.6f); worst rel. diff of the Adam update %.3e"
              % (world, prec, losses[0].item(), g1, losses[1].item(), d1, worst), flush=True)
        assert abs(losses[0].item() - g1) / g1 < 1e-4 and abs(losses[1].item() - d1) / d1 < 1e-4
        assert worst < 0.1, worst
        print("MULTIGPU OK", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
