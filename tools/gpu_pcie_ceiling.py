"""What the host path can carry: every rank copies the bytes one e2e bench step moves -- 1 GiB host -> device and
0.64 GiB device -> host, page-locked memory, 8 MiB pieces on two streams, both directions at once -- with no
compute at all, all ranks together.  The aggregate is the ceiling of `e2e` at that N on that box.
usage: [torchrun --nproc-per-node N] python tools/gpu_pcie_ceiling.py"""
import json, os, time
import torch
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
MIB = 1 << 20
n_in, n_out, piece = 1024 * MIB, 653 * MIB, 8 * MIB
h_in = torch.empty(n_in, dtype=torch.uint8, pin_memory=True); h_in.fill_(7)
h_out = torch.empty(n_out, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n_in, dtype=torch.uint8, device="cuda"); d_out = torch.zeros(n_out, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def step(h2d=True, d2h=True):
    if h2d:
        with torch.cuda.stream(s1):
            for off in range(0, n_in, piece):
                d_in[off:off + piece].copy_(h_in[off:off + piece], non_blocking=True)
    if d2h:
        with torch.cuda.stream(s2):
            for off in range(0, n_out, piece):
                h_out[off:off + piece].copy_(d_out[off:off + piece], non_blocking=True)
    s1.synchronize(); s2.synchronize()

def timed(**kw):
    step(**kw)
    torch.cuda.synchronize()
    if dist is not None: dist.barrier()
    t = time.perf_counter()
    for _ in range(5): step(**kw)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 5
    if dist is not None:
        x = torch.tensor([dt], dtype=torch.float64, device="cuda"); dist.all_reduce(x, op=dist.ReduceOp.MAX); dt = float(x.item())
    return dt

both, up, down = timed(), timed(d2h=False), timed(h2d=False)
if rank == 0:
    print(json.dumps({"n_gpus": world, "ms_per_step_both_directions": round(both * 1e3, 2),
                      "e2e_ceiling_GBps_uncompressed": round(world * n_in / both / 1e9, 2),
                      "h2d_alone_GBps": round(world * n_in / up / 1e9, 2), "d2h_alone_GBps": round(world * n_out / down / 1e9, 2),
                      "per_rank": {"h2d_GBps": round(n_in / up / 1e9, 2), "d2h_GBps": round(n_out / down / 1e9, 2)},
                      "cpus_allowed": len(os.sched_getaffinity(0))}))
if dist is not None: dist.destroy_process_group()
