#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native jdeflate codec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mib M] [--level L]

Metric (BASELINE.json): level-6 deflate GB/s of uncompressed bytes (GB = 1e9 B); the inflate
figure of the same run rides in the "inflate" object of the JSON line.

Workload at N = 1: BASELINE.json configs[1] -- gzip level 6 of 1 GiB of the mixed synthetic
corpus (4 MiB segments cycling TEXT / BINARY / INCOMPRESSIBLE).  One step = one pass of the hot
path over that 1 GiB.  At N > 1 the job is ONE gzip member of N GiB (weak scaling: 1 GiB per rank):
rank g compresses the slice jdeflate_b200.shard.plan() gives it, ends with DEFLT_FLUSH (the last
owner with DEFLT_END), and every step runs the one collective of the path -- an all_gather of the
per-rank {compressed bytes, crc32, length} that fixes the output offsets and the whole-stream CRC
(SURVEY.md 8e); after the timed region rank 0 gathers the parts once and decodes the assembled
member (header + parts + trailer).

  value      device-resident: input already in HBM, output to HBM, through the C ABI
             (deflator_* on device pointers, the gzip CRC-32 of the input, the 10-byte header
             and the 8-byte trailer written around the stream in HBM)
  c3 / c4 / c5   the other BASELINE configs at their stated sizes (see the functions below)
  e2e        the same work through the reference-facing zstrm API with HOST buffers: pinned input,
             zstrm_deflate(8 MiB pieces) -> target callback; every byte crosses PCIe inside the
             timed region
  roofline   the dominant kernel, timed live with CUDA events on its own stream (jdb200_profile)
  cpu_baseline  the compiled reference (oracle/_ref) on the host cores, bounded sample
  --impl reference   only the CPU reference arm, same metric / config

Nothing here reads /root/reference at run time.  oracle/ is used only by the cpu_baseline and
--impl reference legs (as the thing timed there) -- never on the product path.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
import zlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

MIB = 1 << 20
SEG = 4 * MIB                      # corpus segment (tools/corpus.c)
MIXED = 5


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------

def corpus_lib():
    from support import Corpus
    return Corpus()


def fill_parallel(corpus, kind, addr, n, offset, threads=None):
    """Generate [offset, offset+n) of a corpus straight into memory at `addr` (segments in parallel)."""
    threads = threads or min(32, os.cpu_count() or 1)
    jobs = []
    pos = 0
    while pos < n:
        k = min(n - pos, SEG - ((offset + pos) % SEG))
        jobs.append((pos, k))
        pos += k
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda j: corpus.fill_into(kind, addr + j[0], j[1], offset + j[0]), jobs))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
    return 7700.0 * 0.84, "fallback (B200_PROFILING.md: ~84 % of 7.7 TB/s nominal)"


# ---------------------------------------------------------------------------------------------
# CPU reference arm (compiled reference on the host cores)
# ---------------------------------------------------------------------------------------------

def cpu_reference(corpus, op, level, sample_bytes, threads, offset=0):
    """Times oracle/_ref/libjdeflate_ref.so (kind "reference") or, where it is absent, the oracle
    port through the same driver.  Returns (GB/s, seconds, compressed bytes, kind)."""
    base = C.CDLL(str(ROOT / "oracle" / "_ref" / "libjd_cpubase.so"))
    base.jdcb_run.restype = C.c_int
    base.jdcb_run.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int,
                              C.POINTER(C.c_double), C.POINTER(C.c_size_t)]
    ref = ROOT / "oracle" / "_ref" / "libjdeflate_ref.so"
    if not ref.exists():
        raise RuntimeError("oracle/_ref/libjdeflate_ref.so missing (build it in the build container: make -C oracle)")
    buf = C.create_string_buffer(sample_bytes)
    fill_parallel(corpus, MIXED, C.addressof(buf), sample_bytes, offset)
    secs, comp = C.c_double(), C.c_size_t()
    rc = base.jdcb_run(str(ref).encode(), op, level, buf, sample_bytes, threads, C.byref(secs), C.byref(comp))
    if rc != 0:
        raise RuntimeError(f"jdcb_run rc={rc}")
    return sample_bytes / secs.value / 1e9, secs.value, comp.value, "reference"


def cpu_sample_size(cores, total):
    # >= 2 s of wall time: 120 MiB of the mixed corpus per core (whole TEXT/BINARY/INCOMP cycles of
    # 12 MiB; the reference does 50-55 MB/s per core at level 6: 60 MiB per core came to 1.1-1.6 s); the
    # sample may be larger than one GPU step's input -- it is a sample of the same corpus, and a
    # longer one is a steadier denominator
    n = 120 * MIB * max(1, cores)
    return max(12 * MIB, n - n % (12 * MIB))


def run_reference_arm(args, real_stdout=sys.stdout):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    corpus = corpus_lib()
    cores = os.cpu_count() or 1
    total = args.mib * MIB
    sample = cpu_sample_size(cores, total)
    times = []
    comp = 0
    for i in range(args.warmup + args.steps):
        gbs, secs, comp, kind = cpu_reference(corpus, 0, args.level, sample, cores)
        if i >= args.warmup:
            times.append(secs)
    ms = 1e3 * sum(times) / len(times)
    value = sample / (ms / 1e3) / 1e9
    inf_gbs, _, _, _ = cpu_reference(corpus, 1, args.level, sample, cores)
    line = {
        "impl": "reference", "metric": "deflate_level%d_GBps_uncompressed" % args.level, "value": round(value, 4),
        "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args),
        "note": "CPU reference: each step is a bounded sample of the workload",
        "cpu_baseline": {"value": round(value, 4), "unit": "GB/s", "cores": cores, "kind": kind,
                         "sample": f"{sample // MIB} MiB of the mixed corpus, one reference TDeflator per core on contiguous slices"},
        "e2e": {"value": round(value, 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "inflate": {"value": round(inf_gbs, 4), "unit": "GB/s", "note": "reference TInflator per core on its own L%d slices" % args.level},
        "ratio": round(sample / comp, 4), "gpu_launches": 0,
    }
    print(json.dumps(line), file=real_stdout, flush=True)
    return 0


def workload_config(args, extra=None):
    c = {"workload": "gzip level %d of %d MiB mixed synthetic corpus (TEXT/BINARY/INCOMPRESSIBLE, 4 MiB segments) per GPU "
                     "[BASELINE configs[1]]" % (args.level, args.mib),
         "level": args.level, "bytes_per_gpu": args.mib * MIB, "container": "gzip",
         "cache": "every step streams %d MiB of input (> the 126 MB L2) from HBM; no explicit flush" % args.mib}
    if extra:
        c.update(extra)
    return c


def verify_sharded(jd, api, torch, np, dist, rank, world, dev_in, dev_out, produced, state, total_bytes, wbits, log,
                   trailer="crc32"):
    """After the timed region: rank 0 receives every rank's part once (NCCL send / recv into its place
    at the gathered offset), so that the assembled container -- header + parts + trailer -- exists
    in one piece, and decodes it: the whole of it through this library's inflator on the GPU
    (output checksum against the member checksum the all_gather produced), its head through zlib
    on the host, and all of it through zlib when it is at most 2 GiB."""
    t0 = time.time()
    if world > 1:
        sizes = torch.tensor([produced], dtype=torch.int64, device="cuda")
        allsz = [torch.zeros_like(sizes) for _ in range(world)]
        dist.all_gather(allsz, sizes)
        allsz = [int(x.item()) for x in allsz]
        offs = [sum(allsz[:r]) for r in range(world)]
        total = sum(allsz)
        if rank != 0:
            dist.send(dev_out[:produced].contiguous(), dst=0)
            return None
        member = torch.empty(total, dtype=torch.uint8, device="cuda")
        member[:produced].copy_(dev_out[:produced])
        for r in range(1, world):
            dist.recv(member[offs[r]:offs[r] + allsz[r]], src=r)
        torch.cuda.synchronize()
    else:
        offs, total, member = [0], produced, dev_out[:produced]
    head, tail = (10, 8) if wbits == 31 else (2, 4)
    back = torch.empty(total_bytes, dtype=torch.uint8, device="cuda")
    si = jd.inflator()
    si.setsrc(member.data_ptr() + head, total - head - tail)
    si.settgt(back.data_ptr(), total_bytes)
    r = si.inflate(1)
    assert r == api.OK and si.tgtend() == total_bytes and si.srcend() == total - head - tail, (r, si.error, si.tgtend())
    si.close()
    tr = member[total - tail:].cpu().numpy().tobytes()
    if trailer == "crc32":
        got = jd.lib.zstrm_crc32update(0xFFFFFFFF, back.data_ptr(), total_bytes) ^ 0xFFFFFFFF
        assert got == state["member_crc"], "sharded member: decoded CRC-32 differs from the combined CRC"
        assert tr == got.to_bytes(4, "little") + (total_bytes & 0xFFFFFFFF).to_bytes(4, "little"), "gzip trailer"
    else:
        got = jd.lib.zstrm_adler32update(1, back.data_ptr(), total_bytes)
        assert got == state["member_crc"], "sharded stream: decoded Adler-32 differs from the combined Adler-32"
        assert tr == got.to_bytes(4, "big"), "zlib trailer"
    n0 = dev_in.numel()
    assert torch.equal(back[:n0], dev_in), "sharded member: rank 0's slice differs"
    check = min(total_bytes, 64 * MIB)
    hostm = member[: min(total, check + (8 << 20))].cpu().numpy().tobytes()
    zhead = zlib.decompressobj(wbits).decompress(hostm, check)
    assert zhead == back[:check].cpu().numpy().tobytes(), "sharded member: zlib disagrees on the head"
    zfull = False
    if total_bytes <= (2 << 30) and world > 1:
        full = zlib.decompress(member.cpu().numpy().tobytes(), wbits)
        assert len(full) == total_bytes and zlib.crc32(full) == (jd.lib.zstrm_crc32update(0xFFFFFFFF, back.data_ptr(), total_bytes) ^ 0xFFFFFFFF)
        zfull = True
        del full
    log(f"sharded stream verified: {world} parts, {total} compressed bytes, decoded {total_bytes} bytes in {time.time() - t0:.1f}s")
    return {"parts": world, "compressed_bytes": total, "offsets": offs, "decoded_by": ["jdeflate-b200 inflator (whole)",
            "zlib (whole)" if zfull else "zlib (first %d MiB)" % (check >> 20)], "trailer": trailer, "ok": True}


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------

def bind_to_gpu_numa_node(torch, local):
    """One process per GPU: run on (and allocate the page-locked host buffers from) the NUMA node
    the GPU hangs off -- with every rank on node 0 the host staging traffic of the GPUs of the other
    socket crosses the inter-socket link twice.  Best effort (containers may pin the CPU set); returns
    what was done for the JSON line."""
    import ctypes
    out = {"node": None, "cpus": None, "mempolicy": None}
    try:
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, dev)
        node = int(open(path).read().strip())
        if node < 0:
            return out
        out["node"] = node
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if use:
            os.sched_setaffinity(0, use)
            out["cpus"] = len(use)
        # memory of this process (cudaHostAlloc included) preferably from that node: MPOL_PREFERRED = 1
        libc = ctypes.CDLL(None, use_errno=True)
        mask = ctypes.c_ulong(1 << node)
        rc = libc.syscall(238, 1, ctypes.byref(mask), ctypes.c_ulong(64))     # __NR_set_mempolicy on x86-64
        out["mempolicy"] = "preferred" if rc == 0 else "errno %d" % ctypes.get_errno()
    except Exception as ex:                                                     # reported, never required
        out["note"] = repr(ex)
    return out


def _claim_stdout():
    """Everything any library prints to fd 1 during the run (NCCL's version banner, for one) goes
    to stderr; the returned file object is the real stdout, used once for the JSON line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    real_stdout = _claim_stdout()
    try:
        return _main(real_stdout)
    finally:
        real_stdout.flush()


def _main(real_stdout):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mib", type=int, default=1024, help="uncompressed MiB per GPU per step")
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--records", type=int, default=65536, help="distinct zlib JSON records of the batched inflate leg (C3)")
    ap.add_argument("--record-count", type=int, default=1 << 20, help="records per step of the batched inflate leg (C3), all ranks together")
    ap.add_argument("--c4-gib", type=float, default=16.0, help="GiB of LOGS of the sharded zlib level-9 leg (C4), all ranks together")
    ap.add_argument("--no-c4", action="store_true")
    ap.add_argument("--no-c5", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-inflate", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        log("note: the timing rules ask for >= 3 warm-up steps")

    if args.impl == "reference":
        return run_reference_arm(args, real_stdout)

    import numpy as np
    import torch
    from jdeflate_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa_node(torch, local) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    jd = api.load()
    assert jd.lib.jdb200_set_device(local) == 0, jd.lib.jdb200_last_error()
    from jdeflate_b200 import shard
    corpus = corpus_lib()
    chunk_bytes = int(os.environ.get("JDB200_CHUNK_KIB", "512")) << 10
    # ONE gzip member over the whole job: the slice of this rank (weak scaling: --mib per rank)
    total_bytes = world * args.mib * MIB
    sp = shard.plan(total_bytes, world, rank, chunk_bytes)
    n = sp.end - sp.begin
    cap = n + n // 8 + 65536
    GZ_HEAD = bytes([0x1F, 0x8B, 8, 0, 0, 0, 0, 0, 0, 0xFF])

    # ---- workload: this rank's slice of the mixed corpus, generated into pinned host memory ----
    t0 = time.time()
    host_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    fill_parallel(corpus, MIXED, host_in.data_ptr(), n, offset=sp.begin)
    dev_in = host_in.cuda(non_blocking=False)
    dev_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    head_len = len(GZ_HEAD) if rank == 0 else 0
    gz_head = torch.frombuffer(bytearray(GZ_HEAD), dtype=torch.uint8).pin_memory()
    gz_tail = torch.zeros(8, dtype=torch.uint8).pin_memory()
    log(f"[rank {rank}] corpus ready in {time.time() - t0:.1f}s (bytes {sp.begin}..{sp.end} of {total_bytes}, last={sp.last})")

    d = jd.deflator(args.level)
    state = {}

    def step_device():
        """One pass: gzip header (rank 0), raw DEFLATE of the slice, CRC-32 of the slice, the
        all_gather that fixes offsets and the member CRC, the trailer (last rank) -- all in HBM."""
        if head_len:
            dev_out[:head_len].copy_(gz_head, non_blocking=True)
        d.reset()
        d.setsrc(dev_in.data_ptr(), n)
        d.settgt(dev_out.data_ptr() + head_len, cap - head_len - 8)
        r = d.deflate(api.DEFLT_END if sp.last else api.DEFLT_FLUSH)
        assert r == api.OK, (r, d.error, jd.lib.jdb200_last_error())
        crc = jd.lib.zstrm_crc32update(0xFFFFFFFF, dev_in.data_ptr(), n) ^ 0xFFFFFFFF
        produced = head_len + d.tgtend()
        member_crc, member_len, offset, total_comp = crc, n, 0, produced
        if dist is not None:
            # the one collective of the path, every step
            offset, offsets, total_comp, member_crc, member_len = shard.exchange(dist, "cuda", produced + (8 if sp.last else 0), crc, n, "crc32")
        if sp.last:
            gz_tail.numpy()[:] = np.frombuffer(member_crc.to_bytes(4, "little") + (member_len & 0xFFFFFFFF).to_bytes(4, "little"), np.uint8)
            dev_out[produced:produced + 8].copy_(gz_tail, non_blocking=True)
            produced += 8
        state.update(produced=produced, crc=crc, member_crc=member_crc, offset=offset, total_comp=total_comp)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # clocks are sampled from before the warm-up steps (same load) to the end of the timed region:
    # nvidia-smi needs ~1 s to deliver its first sample, longer than a short timed region
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step_device()

    # ---- timed region: device-resident ----------------------------------------------------------
    jd.profile(True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    prof = jd.profile_read()
    jd.profile(False)
    ms_step = ms_total / args.steps
    value = total_bytes / (ms_step / 1e3) / 1e9
    produced = state["produced"]
    launches = sum(v[0] for v in prof.values())

    # correctness of what was timed (not in the timed region)
    assert state["crc"] == zlib.crc32(host_in.numpy()), "CRC-32 mismatch"
    sharded = None
    if dist is None:
        # the gzip member decodes through zlib (first 64 MiB), carries the right trailer, and our
        # own inflator gives the whole input back
        comp = dev_out[:produced].cpu().numpy().tobytes()
        check_n = min(n, 64 * MIB)
        back = zlib.decompressobj(31).decompress(comp, check_n)
        assert back == host_in[:check_n].numpy().tobytes(), "decoded bytes differ from the input"
        assert comp[:10] == GZ_HEAD and comp[-8:-4] == state["crc"].to_bytes(4, "little") and comp[-4:] == (n & 0xFFFFFFFF).to_bytes(4, "little")
        del comp, back
    else:
        sharded = verify_sharded(jd, api, torch, np, dist, rank, world, dev_in, dev_out, produced, state, total_bytes, 31, log)
    raw_produced = produced - head_len - (8 if sp.last else 0)

    # ---- roofline of the dominant kernel ------------------------------------------------------------
    peak, peak_src = measured_peaks()
    dom = max(prof.items(), key=lambda kv: kv[1][1]) if prof else ("none", (0, 0.0))
    dom_name, (dom_launches, dom_ms) = dom
    deflate_kernels = {k: v for k, v in prof.items() if not k.startswith("ck_")}
    # one launch of a deflate-pipeline kernel covers one batch: algorithmic bytes = N (read) + C (written)
    nbatches = max(1, sum(v[0] for k, v in prof.items() if k.startswith("lz_kernel")) or args.steps)
    alg_bytes_per_launch = (n + raw_produced) * args.steps / nbatches
    achieved = alg_bytes_per_launch / (dom_ms / max(dom_launches, 1) / 1e3) / 1e9 if dom_ms else 0.0
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 5), "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": int(alg_bytes_per_launch),
                "avg_launch_ms": round(dom_ms / max(dom_launches, 1), 4),
                "kernel_ms_per_step": {k: round(v[1] / args.steps, 3) for k, v in sorted(prof.items())},
                "kernel_share_of_step": {k: round(v[1] / args.steps / ms_step, 4) for k, v in sorted(prof.items())}}
    traffic_file = ROOT / "profiles" / "traffic.json"
    if traffic_file.exists():
        try:
            roofline["traffic"] = json.loads(traffic_file.read_text()).get(dom_name.split("<")[0])
        except Exception:
            pass

    # ---- checksums of the same 1 GiB (device-resident): HBM-read bound ---------------------------------
    checksum = {}
    try:
        for name, fn, start in (("crc32", jd.lib.zstrm_crc32update, 0xFFFFFFFF), ("adler32", jd.lib.zstrm_adler32update, 1)):
            fn(start, dev_in.data_ptr(), n)
            jd.profile(True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            reps = 10
            for _ in range(reps):
                got = fn(start, dev_in.data_ptr(), n)
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) / reps
            kp = jd.profile_read()
            jd.profile(False)
            kms = sum(v[1] for k, v in kp.items() if k.startswith("ck_")) / reps
            want = zlib.crc32(host_in.numpy()) ^ 0xFFFFFFFF if name == "crc32" else zlib.adler32(host_in.numpy())
            assert got == want, name
            checksum[name] = {"value": round(n / wall / 1e9, 1), "unit": "GB/s", "kernel_GBps": round(n / (kms / 1e3) / 1e9, 1),
                              "frac_of_hbm_peak": round(n / (kms / 1e3) / 1e9 / peak, 4)}
    except Exception as ex:                                           # reported, never required
        checksum = {"note": f"failed: {ex}"}

    # ---- e2e: zstrm API, host buffers -----------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        host_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
        out_base = host_out.data_ptr()
        pos = {"p": 0}

        keep = {"on": True}

        def sink(buf, size, user):
            # the bytes are in host memory (page-locked I/O buffer of the library) when the callback
            # runs; the verification step copies them out, the timed steps only count them (a
            # consumer that streams them on, e.g. to a socket, without another host copy)
            if keep["on"]:
                C.memmove(out_base + pos["p"], buf, size)
            pos["p"] += size
            return size
        cb = api.OFN(sink)
        piece = 8 * MIB
        # one instance, re-armed with zstrm_reset per object (the reference's reuse path,
        # src/zstrm.c:196): buffers and CUDA resources are kept across objects
        z = jd.lib.zstrm_create(api.ZSTRM_DEFLATE | api.ZSTRM_GZIP, args.level, None)
        assert z

        def step_host():
            jd.lib.zstrm_reset(z)
            pos["p"] = 0
            jd.lib.zstrm_settargetfn(z, cb, None)
            base = host_in.data_ptr()
            for off in range(0, n, piece):
                k = min(piece, n - off)
                got = jd.lib.zstrm_deflate(z, base + off, k)
                assert got == k, (got, z.contents.error)
            jd.lib.zstrm_flush(z, 1)
            assert z.contents.error == 0 and z.contents.state == 4
            return z.contents.crc

        crc = step_host()                       # warm-up + the copy that is verified below
        gz_len = pos["p"]
        keep["on"] = False
        step_host()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 5))
        for _ in range(e2e_steps):
            crc2 = step_host()
        barrier()
        secs = (time.perf_counter() - t0) / e2e_steps
        assert crc2 == crc and pos["p"] == gz_len
        jd.lib.zstrm_destroy(z)
        if dist is not None:
            t = torch.tensor([secs], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = float(t.item())
        gz = host_out[:gz_len].numpy().tobytes()
        assert crc == zlib.crc32(host_in.numpy())
        assert gz[-8:-4] == crc.to_bytes(4, "little") and gz[-4:] == (n & 0xFFFFFFFF).to_bytes(4, "little")
        head = zlib.decompressobj(31).decompress(gz, 8 * MIB)
        assert head == host_in[: len(head)].numpy().tobytes(), "e2e: decoded bytes differ from the input"
        e2e = {"value": round(world * n / secs / 1e9, 3), "unit": "GB/s", "h2d_bytes_per_step": n,
               "d2h_bytes_per_step": gz_len, "ms_per_step": round(secs * 1e3, 2), "steps": e2e_steps,
               "api": "zstrm_reset / zstrm_settargetfn / zstrm_deflate(8 MiB pieces from pinned host memory) / zstrm_flush(1) on one reused "
                      "zstrm(DEFLATE|GZIP) instance; the target callback receives every compressed byte in host memory"}
        del gz

    # ---- inflate leg: batched zlib JSON records (BASELINE configs[2], scaled) ----------------------------
    inflate = None
    if not args.no_inflate:
        inflate = bench_inflate(jd, corpus, args, torch, np, barrier, peak, rank, world, dist)

    # ---- the other BASELINE configs at their stated sizes ---------------------------------------------------
    c5 = None
    if not args.no_c5:
        try:
            c5 = bench_c5(jd, args, torch, np, barrier, host_in, n, dist, world, e2e)
        except Exception as ex:                                       # reported, never required
            if dist is not None:
                raise
            c5 = {"note": f"failed: {ex!r}"}
    c4 = None
    if not args.no_c4:
        del dev_out
        torch.cuda.empty_cache()
        c4 = bench_c4(jd, corpus, args, torch, np, barrier, rank, world, dist, chunk_bytes)

    # ---- CPU baseline beside it (rank 0, N = 1 only) ------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        sample = cpu_sample_size(cores, n)
        try:
            gbs, secs, csz, kind = cpu_reference(corpus, 0, args.level, sample, cores)
            one = min(sample, 24 * MIB)
            gbs1, secs1, _, _ = cpu_reference(corpus, 0, args.level, one, 1)
            igbs, _, _, _ = cpu_reference(corpus, 1, args.level, sample, cores)
            cpu = {"value": round(gbs, 4), "unit": "GB/s", "cores": cores, "kind": kind,
                   "sample": f"{sample // MIB} MiB of the same mixed corpus, one reference TDeflator per core on contiguous "
                             f"slices ({secs:.1f} s wall)",
                   "single_thread": {"value": round(gbs1, 4), "sample": f"{one // MIB} MiB"},
                   "inflate": {"value": round(igbs, 4), "unit": "GB/s", "cores": cores},
                   "ratio": round(sample / csz, 4)}
        except Exception as ex:           # the baseline is reported, never required
            cpu = {"value": None, "unit": "GB/s", "cores": cores, "kind": "reference", "sample": f"failed: {ex}"}

    if rank == 0:
        line = {
            "metric": "deflate_level%d_GBps_uncompressed" % args.level, "value": round(value, 3), "unit": "GB/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args),
            "implementation": {"chunk_kib": chunk_bytes >> 10,
                               "collective": "all_gather of 24 B per rank, every step" if world > 1 else "none",
                               "sharding": "one gzip member over %d ranks: DEFLT_FLUSH on all but the last" % world if world > 1 else "none"},
            "ratio": round(n / raw_produced, 4), "compressed_bytes_per_gpu": produced, "sharded": sharded, "numa": numa,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "inflate": inflate, "c4": c4, "c5": c5, "checksum": checksum,
            "gpu_launches": launches, "clocks": clocks,
        }
        print(json.dumps(line), file=real_stdout, flush=True)
    d.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def bench_c4(jd, corpus, args, torch, np, barrier, rank, world, dist, chunk_bytes):
    """BASELINE configs[3]: zlib level 9 of 16 GiB of synthetic logs as ONE zlib stream, strong-scaled
    over the ranks (16/N GiB each): slices from shard.plan() at the codec's chunk size, DEFLT_FLUSH
    on every rank but the last owner, the all_gather of {compressed bytes, Adler-32, length} inside
    every timed step; afterwards rank 0 gathers the parts once and decodes header + parts + trailer."""
    from jdeflate_b200 import api, shard
    LOGS = 1
    total_bytes = int(args.c4_gib * (1 << 30)) // (world * chunk_bytes) * (world * chunk_bytes)
    sp = shard.plan(total_bytes, world, rank, chunk_bytes)
    n = sp.end - sp.begin
    t0 = time.time()
    host = np.empty(n, np.uint8)
    fill_parallel(corpus, LOGS, host.ctypes.data, n, offset=sp.begin)
    dev_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    piece = 256 * MIB
    for off in range(0, n, piece):
        dev_in[off:off + piece].copy_(torch.from_numpy(host[off:off + piece]))
    del host
    cap = n // 2 + n // 8 + (64 << 20)            # logs compress > 4:1; the encoder reports TGTEXHSTD otherwise
    dev_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    ZHEAD = bytes([0x78, 0xDA])
    head_len = 2 if rank == 0 else 0
    zhead = torch.frombuffer(bytearray(ZHEAD), dtype=torch.uint8).pin_memory()
    ztail = torch.zeros(4, dtype=torch.uint8).pin_memory()
    log(f"[rank {rank}] C4: {n / (1 << 30):.2f} GiB of logs ready in {time.time() - t0:.1f}s")
    d = jd.deflator(9)
    state = {}

    def step():
        if head_len:
            dev_out[:head_len].copy_(zhead, non_blocking=True)
        d.reset()
        d.setsrc(dev_in.data_ptr(), n)
        d.settgt(dev_out.data_ptr() + head_len, cap - head_len - 4)
        r = d.deflate(api.DEFLT_END if sp.last else api.DEFLT_FLUSH)
        assert r == api.OK, (r, d.error, jd.lib.jdb200_last_error())
        adler = jd.lib.zstrm_adler32update(1, dev_in.data_ptr(), n)
        produced = head_len + d.tgtend()
        member, length, offset, total_comp = adler, n, 0, produced
        if dist is not None:
            offset, offsets, total_comp, member, length = shard.exchange(dist, "cuda", produced + (4 if sp.last else 0), adler, n, "adler32")
        if sp.last:
            ztail.numpy()[:] = np.frombuffer(member.to_bytes(4, "big"), np.uint8)
            dev_out[produced:produced + 4].copy_(ztail, non_blocking=True)
            produced += 4
        state.update(produced=produced, member_crc=member, offset=offset, total_comp=total_comp)

    step()
    jd.profile(True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    steps = 2
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    prof = jd.profile_read()
    jd.profile(False)
    ver = verify_sharded(jd, api, torch, np, dist, rank, world, dev_in, dev_out, state["produced"], state, total_bytes, 15, log,
                         trailer="adler32")
    d.close()
    if rank != 0:
        return None
    return {"value": round(total_bytes / (ms / 1e3) / 1e9, 3), "unit": "GB/s", "ms_per_step": round(ms, 2), "steps": steps,
            "scaling": "strong", "workload": "zlib level 9 of %.2f GiB of synthetic logs as one stream over %d rank(s) "
            "[BASELINE configs[3]]" % (total_bytes / (1 << 30), world), "bytes": total_bytes, "bytes_per_gpu": n,
            "ratio": round(total_bytes / ver["compressed_bytes"], 4), "sharded": ver,
            "kernel_ms_per_step": {k: round(v[1] / steps, 2) for k, v in sorted(prof.items()) if v[1] > 0.5 * steps}}


def bench_c5(jd, args, torch, np, barrier, host_in, n, dist, world, e2e_l6):
    """BASELINE configs[4]: the zstrm gzip streaming pipeline with 8 MiB callbacks and HOST buffers,
    compress and decompress, level 1 and level 6 (the reference's zstrm_deflate / zstrm_inflate with
    callback I/O, src/zstrm.c:792-958, 1112-1313).  Every byte crosses PCIe inside the timed region.
    GB/s of uncompressed bytes; at N > 1 every rank runs its own pipeline (aggregate, max time)."""
    import ctypes as C
    from jdeflate_b200 import api
    piece = 8 * MIB
    cap = n + n // 8 + 65536
    host_gz = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
    host_back = torch.empty(piece, dtype=torch.uint8, pin_memory=True)
    gz_base, in_base = host_gz.data_ptr(), host_in.data_ptr()
    want_crc = zlib.crc32(host_in.numpy())
    out = {}

    def agg(secs):
        if dist is not None:
            t = torch.tensor([secs], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            secs = float(t.item())
        return round(world * n / secs / 1e9, 3)

    for level in (1, 6):
        pos = {"p": 0}

        def sink(buf, size, user):
            C.memmove(gz_base + pos["p"], buf, size)
            pos["p"] += size
            return size
        ocb = api.OFN(sink)
        z = jd.lib.zstrm_create(api.ZSTRM_DEFLATE | api.ZSTRM_GZIP, level, None)
        assert z

        def compress():
            jd.lib.zstrm_reset(z)
            pos["p"] = 0
            jd.lib.zstrm_settargetfn(z, ocb, None)
            for off in range(0, n, piece):
                k = min(piece, n - off)
                assert jd.lib.zstrm_deflate(z, in_base + off, k) == k, z.contents.error
            jd.lib.zstrm_flush(z, 1)
            assert z.contents.error == 0

        compress()
        barrier()
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            compress()
        barrier()
        csecs = (time.perf_counter() - t0) / reps
        gz_len = pos["p"]
        jd.lib.zstrm_destroy(z)

        rd = {"p": 0}

        def source(buf, size, user):
            k = min(size, piece, gz_len - rd["p"])
            C.memmove(buf, gz_base + rd["p"], k)
            rd["p"] += k
            return k
        icb = api.IFN(source)
        zi = jd.lib.zstrm_create(api.ZSTRM_INFLATE | api.ZSTRM_GZIP, 0, None)
        assert zi

        def decompress(check):
            jd.lib.zstrm_reset(zi)
            rd["p"] = 0
            jd.lib.zstrm_setsourcefn(zi, icb, None)
            total, crc = 0, 0
            while True:
                got = jd.lib.zstrm_inflate(zi, host_back.data_ptr(), piece)
                if got <= 0:
                    break
                if check:
                    crc = zlib.crc32(host_back.numpy()[:got], crc)
                total += got
            assert zi.contents.error == 0 and total == n, (zi.contents.error, total)
            if check:
                assert crc == want_crc, "C5: decompressed bytes differ"

        decompress(True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            decompress(False)
        barrier()
        dsecs = (time.perf_counter() - t0) / reps
        jd.lib.zstrm_destroy(zi)
        out["level%d" % level] = {"deflate": agg(csecs), "inflate": agg(dsecs), "unit": "GB/s", "ratio": round(n / gz_len, 4),
                                  "h2d_d2h_bytes_per_step": {"deflate": [n, gz_len], "inflate": [gz_len, n]}}
    out["workload"] = ("zstrm gzip streaming compress + decompress of %d MiB mixed corpus per GPU, host (pinned) buffers, 8 MiB "
                       "pieces through the target / source callbacks [BASELINE configs[4]]" % (n // MIB))
    if e2e_l6:
        out["level6"]["deflate_headline_e2e"] = e2e_l6["value"]
    return out


def bench_inflate(jd, corpus, args, torch, np, barrier, peak, rank=0, world=1, dist=None):
    """BASELINE configs[2] at its stated size: batched inflate of 1 Mi independent zlib JSON records
    (4-64 KiB, zlib level 6 = third-party streams, 65 536 distinct ones in a random permutation: the
    compressed working set is far larger than the L2), one warp per record, device-resident
    buffers.  At N > 1 the records are dealt out to the ranks."""
    from jdeflate_b200 import api
    nd = max(1, args.records // world)
    count = max(nd, args.record_count // world)
    t0 = time.time()
    workers = min(32, os.cpu_count() or 1)
    with ThreadPoolExecutor(max_workers=workers) as ex:
        recs = list(ex.map(lambda i: corpus.json_record(i + rank * nd), range(nd), chunksize=64))
        comp = list(ex.map(lambda r: zlib.compress(r, 6), recs, chunksize=64))
    tile = count // nd
    perm = np.random.RandomState(1 + rank).permutation(count) % nd
    clen = np.array([len(x) for x in comp], np.uint64)
    rlen = np.array([len(x) for x in recs], np.uint64)
    src_off = np.zeros(nd + 1, np.uint64)
    src_off[1:] = np.cumsum(clen)
    items = np.zeros((count, 4), np.uint64)
    items[:, 0] = src_off[perm]
    items[:, 2] = clen[perm]
    items[:, 3] = rlen[perm]
    items[1:, 1] = np.cumsum(rlen[perm])[:-1]
    total_out = int(rlen[perm].sum())
    total_in = int(clen[perm].sum())
    src = torch.from_numpy(np.frombuffer(b"".join(comp), np.uint8).copy()).cuda()
    out = torch.empty(total_out + 64, dtype=torch.uint8, device="cuda")
    ditems = torch.from_numpy(items.view(np.int64)).cuda()
    dres = torch.zeros((count, 4), dtype=torch.int64, device="cuda")
    log(f"inflate leg: {nd} distinct records x{tile}, {total_out / 1e9:.2f} GB out, prep {time.time() - t0:.1f}s")

    def step():
        rc = jd.lib.jdb200_inflate_batch(src.data_ptr(), out.data_ptr(), ditems.data_ptr(), dres.data_ptr(), count, 1)
        assert rc == 0, jd.lib.jdb200_last_error()

    for _ in range(3):
        step()
    jd.profile(True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    steps = min(max(3, args.steps), 5)
    for _ in range(steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / steps
    job_out, job_in, job_count = total_out, total_in, count
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        t = torch.tensor([total_out, total_in, count], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        job_out, job_in, job_count = (int(x) for x in t.tolist())
    prof = jd.profile_read()
    jd.profile(False)
    res = dres.cpu().numpy().view(np.uint32).reshape(count, 8)
    assert int((res[:, 0] != 0).sum()) == 0 and int((res[:, 2] != 0).sum()) == 0, "inflate batch reported errors"
    # every record's Adler-32 trailer was checked by the kernel (zerror); bytes of a sample, too
    for k in range(0, count, max(1, count // 256)):
        o, ln = int(items[k, 1]), int(items[k, 3])
        assert out[o:o + ln].cpu().numpy().tobytes() == recs[perm[k]], "inflate output mismatch"
    # the compress-side mirror: the same records, each into its own zlib stream
    batch = None
    try:
        ln = rlen[perm]
        caps = ln + ln // np.uint64(64) + np.uint64(80)
        roff = np.zeros(nd + 1, np.uint64)
        roff[1:] = np.cumsum(rlen)
        it2 = np.zeros((count, 4), np.uint64)
        it2[:, 0] = roff[perm]
        it2[1:, 1] = np.cumsum(caps)[:-1]
        it2[:, 2] = ln
        it2[:, 3] = caps
        rsrc = torch.from_numpy(np.frombuffer(b"".join(recs), np.uint8).copy()).cuda()
        cout = torch.empty(int(caps.sum()) + 64, dtype=torch.uint8, device="cuda")
        d2 = torch.from_numpy(it2.view(np.int64)).cuda()
        r2 = torch.zeros((count, 4), dtype=torch.int64, device="cuda")

        def cstep():
            rc = jd.deflate_batch(rsrc.data_ptr(), cout.data_ptr(), d2.data_ptr(), r2.data_ptr(), count, api.JDB200_ZLIB, args.level)
            assert rc == 0, jd.lib.jdb200_last_error()

        cstep()
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(2):
            cstep()
        c1.record()
        barrier()
        cms = c0.elapsed_time(c1) / 2
        rr = r2.cpu().numpy()
        assert not (rr.view(np.uint32).reshape(count, 8)[:, 0]).any(), "deflate batch reported errors"
        used = rr.view(np.uint64).reshape(count, 4)[:, 3]
        for k in range(0, count, max(1, count // 256)):
            o, n = int(it2[k, 1]), int(used[k])
            assert zlib.decompress(cout[o:o + n].cpu().numpy().tobytes()) == recs[perm[k]], "deflate batch stream mismatch"
        batch = {"value": round(total_out / (cms / 1e3) / 1e9, 3), "unit": "GB/s", "ms_per_step": round(cms, 3),
                 "records_per_s": round(count / (cms / 1e3), 1), "ratio": round(total_out / float(used.sum()), 4),
                 "zlib_ratio": round(total_out / float(total_in), 4),
                 "api": f"jdb200_deflate_batch(JDB200_ZLIB, level {args.level}) on the same {count} records, device buffers"}
        del rsrc, cout, d2, r2
    except Exception as ex:                                           # reported, never required
        batch = {"value": None, "note": f"failed: {ex}"}
    # one large stream of OURS through the plain inflator (chunk-parallel decode)
    own = None
    try:
        nown = args.mib * MIB
        raw = torch.empty(nown, dtype=torch.uint8, pin_memory=True)
        fill_parallel(corpus, MIXED, raw.data_ptr(), nown, offset=0)
        draw = raw.cuda()
        dcomp = torch.empty(nown + nown // 8 + 65536, dtype=torch.uint8, device="cuda")
        de = jd.deflator(args.level)
        de.setsrc(draw.data_ptr(), nown)
        de.settgt(dcomp.data_ptr(), dcomp.numel())
        assert de.deflate(api.DEFLT_END) == api.OK
        clen_own = de.tgtend()
        de.close()
        dback = torch.empty(nown, dtype=torch.uint8, device="cuda")
        best = None
        for _ in range(3):
            si = jd.inflator()
            si.setsrc(dcomp.data_ptr(), clen_own)
            si.settgt(dback.data_ptr(), nown)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rr = si.inflate(1)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            assert rr == api.OK and si.tgtend() == nown and si.srcend() == clen_own
            si.close()
            best = dt if best is None else min(best, dt)
        assert torch.equal(dback, draw)
        own = {"value": round(nown / best / 1e9, 3), "unit": "GB/s", "bytes_out": nown,
               "api": "inflator_inflate(final) on one raw stream produced by this library's deflator (device buffers)"}
        del draw, dcomp, dback
    except Exception as ex:                                           # reported, never required
        own = {"value": None, "note": f"failed: {ex}"}
    # ONE third-party stream (zlib level 6, no sync markers) through the plain inflator: nothing to
    # parallelise over, the sequential decoder (the reference's decodefast on one core: 0.25-0.33 GB/s)
    single = None
    try:
        nsingle = 64 * MIB
        raw1 = np.empty(nsingle, np.uint8)
        fill_parallel(corpus, MIXED, raw1.ctypes.data, nsingle, offset=0)
        z1 = zlib.compress(raw1.tobytes(), 6)[2:-4]
        dz1 = torch.from_numpy(np.frombuffer(z1, np.uint8).copy()).cuda()
        dback1 = torch.empty(nsingle, dtype=torch.uint8, device="cuda")
        best = None
        for _ in range(2):
            si = jd.inflator()
            si.setsrc(dz1.data_ptr(), len(z1))
            si.settgt(dback1.data_ptr(), nsingle)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rr = si.inflate(1)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            assert rr == api.OK and si.tgtend() == nsingle and si.srcend() == len(z1)
            si.close()
            best = dt if best is None else min(best, dt)
        assert dback1.cpu().numpy().tobytes() == raw1.tobytes()
        single = {"value": round(nsingle / best / 1e9, 4), "unit": "GB/s", "bytes_out": nsingle,
                  "api": "inflator_inflate(final) on one raw zlib level-6 stream of the mixed corpus (device buffers)"}
        del dz1, dback1
    except Exception as ex:                                           # reported, never required
        single = {"value": None, "note": f"failed: {ex}"}
    kms = prof.get("inflate_batch_kernel", (steps, ms * steps))
    kavg = kms[1] / max(kms[0], 1)
    ach = (total_in + total_out) / (kavg / 1e3) / 1e9
    return {"value": round(job_out / (ms / 1e3) / 1e9, 3), "unit": "GB/s", "ms_per_step": round(ms, 3),
            "workload": f"batched inflate of {job_count} zlib level-6 JSON records (4-64 KiB, {nd * world} distinct, random permutation"
                        f"{', dealt out to %d ranks' % world if world > 1 else ''}) [BASELINE configs[2]]",
            "records": job_count, "distinct": nd * world, "bytes_out": job_out, "bytes_in": job_in,
            "distinct_compressed_bytes": int(clen.sum()) * world, "records_per_s": round(job_count / (ms / 1e3), 1),
            "own_stream": own, "single_third_party": single, "deflate_batch": batch,
            "roofline": {"bound": "hbm", "kernel": "inflate_batch_kernel", "achieved": round(ach, 2), "peak": peak,
                         "unit": "GB/s", "frac": round(ach / peak, 5), "avg_launch_ms": round(kavg, 4)}}


if __name__ == "__main__":
    sys.exit(main())
