#!/usr/bin/env python3
"""bench.py -- MPEG-1 720p frames/sec (batched) on N B200s, with the roofline of the fused kernel.

One "step" = one pass of the hot path over one batch of synthetic pre-parsed input: per GPU 256
independent 720p streams each decode one dense-P picture (fused motion compensation + 8x8 IDCT +
residual add, one launch) and convert the decoded frame YCbCr->RGBA (one launch).  This is
BASELINE.json configs[2] ("batch 256 synthetic 720p streams on 1 B200"); with --gpus N every rank
holds its own 256 streams (configs[4] layout: stream-parallel, no data-path collective, weak scaling).

  value     whole-job frames/s with the packed records already resident in HBM (device-timed)
  parity    BEFORE any timing the benchmarked configuration proves itself: every stream of the batch is decoded through
            the device-pointer path AND through the host-pointer (end-to-end) paths and compared byte for byte, planes
            and RGBA, with the CPU oracle run on the same records; a mismatch makes the run exit non-zero
  e2e       the same step through the host-pointer C-ABI: pinned host records -> H2D -> kernels -> D2H of the decoded
            Y/Cb/Cr planes into pinned host memory, all inside the timed region.  The coefficients travel in the
            variable-width form the product's own host parser emits (mpegb200_video_parser_set_vlen; the synthetic
            batches are put into that form once, untimed, like they are generated untimed).  e2e_int16 = the same with
            plain int16[64] blocks; e2e_rgba_back = the RGBA frames read back instead of the planes.
  roofline  decode call (plan pre-pass + fused MC+IDCT+add kernel): algorithmic bytes (1552 B per dense-P
            macroblock) / mean launch duration measured with CUDA events on the launching stream, against the measured
            HBM peak of MEASURED_PEAKS.json
  picture_steps
            the other picture steps of the configuration (natural P, natural B, I-only, dense P with +-64 pixel vectors),
            each parity-checked against the oracle and timed like the headline step, fraction of the roofline on ITS bytes
  audio     BASELINE configs[3]: 1024 MP2 streams x 8 frames per launch, both window modes, parity-checked
  bitstream compressed 720p elementary streams in, frames in HBM out: the lock-step batch with the host parser
            against the same batch with the slices parsed on the device (vlc_parse_kernel), every stream's frames hashed
            against the CPU oracle's decode of its bitstream
  sustained the headline step looped for --sustain-seconds with clocks sampled (burst vs sustained)
  transfers H2D / D2H bandwidth of this rank's PCIe link and the host's memcpy bandwidth (explains e2e at N > 1)
  gathered  (N > 1) the frames of every step gathered to rank 0 over NVLink inside the timed region, double-buffered
  cpu_baseline / --impl reference
            the CPU restatement of the reference's Go path (oracle/, validated against the reference's
            golden hashes; no Go toolchain exists in this image) on the host cores, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))   # workload generator and oracle bindings (test / bench infrastructure)

METRIC = "mpeg1_720p_frames_per_sec_batched"
UNIT = "frames/s"
STREAMS_PER_GPU = 256
CONFIG_ID = 3  # SURVEY 8d numbering of "batch 256 synthetic 720p streams"
AUDIO_STREAMS, AUDIO_FRAMES, AUDIO_BYTES_PER_FRAME = 1024, 8, 20352   # configs[3]; SURVEY 8d: 9216 in + 9216 out + (8 KiB state in + out) / 8 frames


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=STREAMS_PER_GPU, help="streams per GPU (default: the BASELINE config)")
    ap.add_argument("--mode", default="dense", choices=["dense", "natural"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run parity check (profiling runs only)")
    ap.add_argument("--no-extras", action="store_true", help="skip the steps / audio / sustained / transfers sub-records")
    ap.add_argument("--sustain-seconds", type=float, default=3.0)
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="target CPU work for cpu_baseline")
    ap.add_argument("--gather-mode", default="p2p", choices=["p2p", "gather"], help="N > 1: how the frames reach rank 0")
    return ap.parse_args()


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic():
    """DRAM bytes per launch of the fused kernel from the committed ncu --set full capture, if any."""
    p = ROOT / "profiles" / "fused_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text())
        except Exception:
            return None
    return None


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled from a thread (about one sample
    per millisecond -- the timed region of the default run is only ~15 ms), nvidia-smi -lms as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml = index, [], None, None
        self.samples, self.power, self.reason_bits, self._stop = [], [], 0, threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.samples.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                self.reason_bits |= int(get_reasons(self.h))
                if len(self.samples) % 64 == 0:
                    self.power.append(n.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.0005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml:
            self._stop.set()
            self.t.join(timeout=2)
            n, bits, reasons = self.nvml, self.reason_bits, []
            for name, const in [("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                                ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                                ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                                ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap")]:
                if bits & int(getattr(n, const, 0)):
                    reasons.append(name)
            return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                    "reasons": reasons, "samples": len(self.samples), "source": "nvml",
                    "power_w_max": max(self.power) if self.power else None}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                for n, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def parse_cpulist(text):
    """'0-3,8,10-11' -> {0, 1, 2, 3, 8, 10, 11} (the format of /sys/.../local_cpulist)."""
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_cpus(device_index, log, sysfs_root="/sys/bus/pci/devices"):
    """Multi-rank runs: keep this rank on the CPUs that are local to its GPU (sysfs local_cpulist of the PCI device), so
    that the pinned staging buffers it allocates afterwards land on that NUMA node (first touch) and eight ranks do not all
    pull their PCIe traffic through one socket's memory.  Best effort: any failure leaves the affinity as it was."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        path = "%s/%04x:%02x:%02x.0/local_cpulist" % (sysfs_root, pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(path) as f:
            local = parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        want = local & allowed
        if len(want) >= 4 and want != allowed:   # never squeeze the launching thread and the clock sampler onto a core or two
            os.sched_setaffinity(0, want)
            log(f"rank bound to {len(want)} of {len(allowed)} CPUs local to GPU {device_index}")
    except Exception as e:  # no sysfs entry, no NUMA information, a cpuset that forbids it ...
        log(f"CPU binding skipped: {e!r}")


def build_batch(streams, first_stream_id, mode, log=None, pic_type=None, mv_range=32, seed_offset=0, with_refs=True):
    """Packed records of one picture step for `streams` 720p streams (one picture each)."""
    import workload as wl
    g = wl.HD720
    pic_type = wl.PIC_P if pic_type is None else pic_type
    per, refs = [], []
    for s in range(streams):
        rng = wl.stream_rng(CONFIG_ID, seed_offset + first_stream_id + s)
        if with_refs:
            refs.append(wl.random_reference_frame(rng, g))
        per.append(wl.make_picture(rng, g, pic_type, mode, mv_range=mv_range))
        if log and (s + 1) % 64 == 0:
            log(f"generated {s + 1}/{streams} streams")
    pics, mbs, coeffs = wl.batch_pictures(per, list(range(streams)), pic_type, [(0, 1, 2)] * streams)
    return g, refs, pics, mbs, coeffs


def rotation_variants(pics):
    """The three (dst, fwd, bwd) assignments a chain of P pictures cycles through (video.go:406-433)."""
    import workload as wl
    rot, out = wl.BufferRotation(), []
    for _ in range(3):
        dst, fwd, bwd = rot.begin(wl.PIC_P)
        p = pics.copy()
        p["dst_buf"], p["fwd_buf"], p["bwd_buf"] = dst, fwd, bwd
        out.append((p, dst))
        rot.end(wl.PIC_P)
    return out


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restatement of the reference's Go path) on the host cores
# ------------------------------------------------------------------------------------------------
class CpuArm:
    """One picture per host thread (OpenMP inside the oracle); the RGBA targets are allocated and touched once, outside every
    timed loop, and no Python thread pool is built inside one."""

    def __init__(self, n_pictures, mode):
        import oracle_lib as ol
        self.ol = ol
        ol.lib().orc_use_swar_mc(1)   # the 8-bytes-per-step copyMacroblock of video_noasm.go, faster than the per-pixel form
        g, refs, self.pics, self.mbs, self.coeffs = build_batch(n_pictures, 0, mode)
        self.fs = ol.FrameSet(n_pictures, g.width, g.height)
        for s in range(n_pictures):
            self.fs.whole(s, 1)[:] = refs[s]
        self.out = np.zeros((n_pictures, g.height, g.width, 4), np.uint8)
        self.ids = np.arange(n_pictures, dtype=np.int32)
        self.n = n_pictures

    def step(self, threads, with_rgba=True):
        t0 = time.perf_counter()
        rc = self.fs.exec_pictures(self.pics, self.mbs, self.coeffs, threads=threads)
        assert rc == 0
        if with_rgba:
            self.fs.rgba_batch(self.ids, self.pics["dst_buf"], self.out, threads=threads)
        return time.perf_counter() - t0

    def best_thread_count(self, n):
        """All hardware threads are not always the fastest (SMT siblings share the integer units): give the CPU
        arm whichever of n and n/2 threads is faster on one trial step."""
        cands = [n] if n < 4 else [n, n // 2]
        timing = {t: min(self.step(t) for _ in range(2)) for t in cands}
        return min(timing, key=timing.get)


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: rank 0 only; each step = one picture per host thread of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = host_threads()
    arm = CpuArm(n, args.mode)
    threads = arm.best_thread_count(n)
    for _ in range(args.warmup):
        arm.step(threads)
    t = sum(arm.step(threads) for _ in range(args.steps))
    fps = arm.n * args.steps / t
    sample = f"{arm.n} dense-P 720p pictures per step (one per host thread) + RGBA, {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": f"720p {args.mode}-P fused MC+IDCT+add + YCbCr->RGBA, CPU restatement of the reference Go path "
                               "(Go toolchain absent), sample of the 256-stream batch", "pictures_per_step": arm.n},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample + f" ({threads} of {n} hardware threads: the faster of n and n/2)"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def cpu_baseline(args, log):
    n = host_threads()
    arm = CpuArm(n, args.mode)
    threads = arm.best_thread_count(n)
    t1 = arm.step(threads)  # warm-up / calibration
    rounds = int(max(2, min(2000, args.cpu_seconds / max(t1, 1e-3))))
    t = sum(arm.step(threads) for _ in range(rounds))
    fps = arm.n * rounds / t
    log(f"cpu baseline: {fps:.1f} frames/s on {threads} threads ({rounds} rounds, {t:.1f} s)")
    return {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{rounds} rounds x {arm.n} dense-P 720p pictures (fused-kernel-equivalent replay + RGBA), "
                      f"{t:.1f} s wall, OpenMP over pictures, 8-byte-SWAR copyMacroblock"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def fail_parity(payload):
    print(json.dumps(payload))
    sys.stdout.flush()
    raise SystemExit(3)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1 (one rank per GPU)")
        args.gpus = world

    def log(msg):
        if rank == 0:
            print(f"[bench] {msg}", file=sys.stderr, flush=True)

    import torch
    import mpeg_b200  # fails loudly if libmpegb200.so is missing
    import workload as wl

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if world > 1:
        bind_to_gpu_cpus(local_rank, log)
    S = args.streams
    t0 = time.time()
    from mpeg_b200.sharding import stream_range
    g, refs, pics, mbs, coeffs = build_batch(S, stream_range(rank, world, S)[0], args.mode, log)
    log(f"workload built in {time.time() - t0:.1f} s: {S} streams, {len(mbs)} macroblocks, {len(coeffs)} blocks")
    alg_total, alg_read = wl.algorithmic_bytes(mbs, len(coeffs))
    peak, peak_src = measured_peak()

    stream = torch.cuda.Stream()
    ctx = mpeg_b200.Context(device=local_rank, max_streams=S)
    ctx.set_stream(stream.cuda_stream)
    for s in range(S):
        ctx.video_open(s, g.width, g.height)
        ctx.video_write_frame(s, 1, refs[s])
        ctx.video_write_frame(s, 0, refs[(s + 1) % S])  # every buffer holds plausible pixels from the start
        ctx.video_write_frame(s, 2, refs[(s + 2) % S])
    variants = rotation_variants(pics)
    ctx.video_validate(variants[0][0], mbs, len(coeffs))

    def pinned(a):
        t = torch.empty(a.nbytes, dtype=torch.uint8, pin_memory=True)
        t.numpy()[:] = a.view(np.uint8).reshape(-1)
        return t

    h_mbs, h_coeffs = pinned(mbs), pinned(coeffs)
    # Transfer form of the coefficients on the end-to-end path: the variable-width groups that the product's host parser emits
    # straight from its zig-zag walk (about 49 B per dense block instead of 128).  The synthetic batch is brought into that
    # form here, once and untimed -- it is part of generating the input, like drawing the levels is.
    h_vlen = None
    if not args.no_e2e:
        h_vlen = tuple(pinned(a) for a in ctx.pack_coeffs_vlen(coeffs))
    h_pics = [pinned(p) for p, _ in variants]
    d_mbs, d_coeffs = h_mbs.cuda(), h_coeffs.cuda()
    d_pics = [p.cuda() for p in h_pics]
    rgba_stride = g.width * g.height * 4
    d_rgba = torch.empty(S * rgba_stride, dtype=torch.uint8, device="cuda")
    h_planes = torch.empty(S * g.picture_bytes, dtype=torch.uint8, pin_memory=True)
    ids = np.arange(S, dtype=np.int32)
    torch.cuda.synchronize()

    def step_dev(k, ev=None):
        p, dst = variants[k % 3]
        with torch.cuda.stream(stream):
            if ev:
                ev[0].record(stream)
            ctx.video_decode_pictures_dev(len(p), d_pics[k % 3].data_ptr(), len(mbs), d_mbs.data_ptr(), len(coeffs), d_coeffs.data_ptr())
            if ev:
                ev[1].record(stream)
            ctx.video_rgba_batch_dev(ids, np.full(S, dst, np.uint8), d_rgba.data_ptr(), rgba_stride)

    def decode_host(k, form):
        p, dst = variants[k % 3]
        L = ctx.L
        if form == "vlen":
            rc = L.mpegb200_video_decode_pictures_vlen(ctx.h, len(p), h_pics[k % 3].data_ptr(), len(mbs), h_mbs.data_ptr(), len(coeffs),
                                                       h_vlen[0].data_ptr(), h_vlen[1].data_ptr(), h_vlen[2].data_ptr(), h_vlen[2].numel())
        else:
            rc = L.mpegb200_video_decode_pictures(ctx.h, len(p), h_pics[k % 3].data_ptr(), len(mbs), h_mbs.data_ptr(), len(coeffs), h_coeffs.data_ptr())
        ctx._ck(rc)
        return dst

    def step_e2e(k, form="vlen"):
        dst = decode_host(k, form)
        ctx.video_rgba_batch_dev(ids, np.full(S, dst, np.uint8), d_rgba.data_ptr(), rgba_stride)
        ctx.video_read_pictures(ids, np.full(S, dst, np.uint8), h_planes.data_ptr(), g.picture_bytes)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, steps, with_kernel_events=False):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)] if with_kernel_events else None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record(stream)
        for k in range(steps):
            fn(k, evs[k]) if with_kernel_events else fn(k)
        b.record(stream)
        barrier()
        ms = a.elapsed_time(b)
        kms = [x.elapsed_time(y) for x, y in evs] if evs else None
        return max_over_ranks(ms), kms

    def timed_e2e(fn, steps):
        """The end-to-end steps finish on the library's read-back stream: the closing event is recorded after the compute
        stream has been made to wait for the last read-backs (mpegb200_join_readbacks), so the region holds every copy."""
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ctx.sync()
        a.record(stream)
        for k in range(steps):
            fn(k)
        ctx.join_readbacks()
        b.record(stream)
        barrier()
        ctx.sync()
        return max_over_ranks(a.elapsed_time(b))

    # ---- the benchmarked configuration proves itself (outside every timed region)
    parity = None
    if not args.no_parity:
        parity = check_parity(args, ctx, g, S, refs, variants, mbs, coeffs, step_dev, decode_host, d_rgba, h_planes, ids, log)
        if dist is not None:
            t = torch.tensor([1 if parity["ok"] else 0], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            parity["ok_all_ranks"] = bool(t.item())
        if not parity["ok"]:
            fail_parity({"metric": METRIC, "parity": parity, "error": "GPU output differs from the CPU oracle"})
    del refs

    # ---- device-resident run (value + roofline)
    for k in range(args.warmup):
        step_dev(k)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    ms, kernel_ms = timed(step_dev, args.steps, with_kernel_events=True)
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    frames = S * world * args.steps
    value = frames / (ms * 1e-3)
    fused_ms = float(np.mean(kernel_ms))
    achieved = alg_total / (fused_ms * 1e-3) / 1e9
    traffic = recorded_traffic()
    # the two kernels of the decode call apart: a second pass with the library's own events between them
    # (mpegb200_set_kernel_timing), kept out of the timed region above because an event record between two
    # kernels costs a little stream time
    ctx.set_kernel_timing(True)
    for k in range(args.steps):
        step_dev(k)
    plan_each, fused_each = ctx.kernel_times()
    ctx.set_kernel_timing(False)
    if len(fused_each) == 0:
        raise SystemExit("the TMA kernel did not run (generic fallback selected?)")
    plan_only_ms, fused_only_ms = float(np.mean(plan_each)), float(np.mean(fused_each))

    # ---- end-to-end runs (host buffers, copies inside the timed region)
    e2e = e2e_int16 = e2e_rgba = None
    if not args.no_e2e:
        pic_bytes = int(h_pics[0].numel())
        for k in range(max(1, args.warmup)):
            step_e2e(k)
        e_steps = max(3, args.steps)  # same K as the device-resident run: the fill and drain of the three-stream pipeline are part of it
        ems = timed_e2e(step_e2e, e_steps)
        e2e = {"value": S * world * e_steps / (ems * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(h_mbs.numel() + pic_bytes + sum(t.numel() for t in h_vlen)),
               "d2h_bytes_per_step": int(h_planes.numel()), "steps": e_steps, "ms_per_step": ems / e_steps,
               "returns": "the decoded Y/Cb/Cr planes of every stream (what Video.Decode() hands out); the RGBA conversion runs on the "
                          "device and its output stays in HBM (Frame.RGBA() is a separate, on-demand call in the reference) -- see e2e_rgba_back",
               "coefficient_form": "vlen: the variable-width groups the host parser emits directly (mpegb200_video_parser_set_vlen), no conversion pass",
               "path": "mpegb200_video_decode_pictures_vlen (pinned host records) + rgba_batch_dev + read_pictures_host (pinned)"}
        # the same with plain int16[64] blocks (north_star's literal input form): PCIe moves 2.5x the bytes
        i_steps = max(3, min(args.steps, 10))
        for k in range(2):
            step_e2e(k, "int16")
        ims = timed_e2e(lambda k: step_e2e(k, "int16"), i_steps)
        e2e_int16 = {"value": S * world * i_steps / (ims * 1e-3), "unit": UNIT,
                     "h2d_bytes_per_step": int(h_mbs.numel() + pic_bytes + h_coeffs.numel()), "d2h_bytes_per_step": int(h_planes.numel()),
                     "steps": i_steps, "ms_per_step": ims / i_steps, "path": "mpegb200_video_decode_pictures (int16 blocks) + rgba_batch_dev + read_pictures_host"}
        if not args.no_extras:
            # and with the RGBA frames coming back instead of the planes
            h_rgba = torch.empty(S * rgba_stride, dtype=torch.uint8, pin_memory=True)
            dl = torch.cuda.Stream()
            back = [None]

            def step_rgba_back(k):
                dst = decode_host(k, "vlen")
                if back[0] is not None:
                    stream.wait_event(back[0])   # the conversion must not overwrite d_rgba under the previous copy
                ctx.video_rgba_batch_dev(ids, np.full(S, dst, np.uint8), d_rgba.data_ptr(), rgba_stride)
                done = torch.cuda.Event()
                done.record(stream)              # the library launches on the stream it was given
                dl.wait_event(done)
                with torch.cuda.stream(dl):
                    h_rgba.copy_(d_rgba, non_blocking=True)
                back[0] = torch.cuda.Event()
                back[0].record(dl)

            def timed_rgba(steps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()
                ctx.sync()
                a.record(stream)
                for k in range(steps):
                    step_rgba_back(k)
                stream.wait_event(back[0])
                b.record(stream)
                barrier()
                return max_over_ranks(a.elapsed_time(b))

            r_steps = 5
            step_rgba_back(0)
            rms = timed_rgba(r_steps)
            e2e_rgba = {"value": S * world * r_steps / (rms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": e2e["h2d_bytes_per_step"],
                        "d2h_bytes_per_step": int(h_rgba.numel()), "steps": r_steps, "ms_per_step": rms / r_steps,
                        "path": "decode_pictures_vlen + rgba_batch_dev + D2H of the RGBA frames (pinned)"}
            del h_rgba

    extras = {}
    if not args.no_extras:
        extras["transfers"] = transfer_probe(torch)
    if not args.no_extras and world == 1:
        extras["sustained"] = sustained_leg(args, torch, stream, step_dev, local_rank, S, alg_total, peak)
        extras["picture_steps"] = picture_steps(torch, ctx, stream, g, S, peak, log)
        extras["audio"] = audio_leg(torch, local_rank, peak, log)
        extras["bitstream"] = bitstream_leg(args, local_rank, log)
    elif not args.no_extras:
        # N > 1: every rank decodes its own 256 streams from their bitstreams (resident path only), barrier + max over ranks
        extras["bitstream"] = bitstream_leg(args, local_rank, log, paths=("device_vlc_resident", "device_vlc_resident_to_host"), barrier=barrier,
                                            max_over_ranks=max_over_ranks, world=world)

    # ---- frames gathered to rank 0 over NVLink (the only collective of the path)
    gather = gathered = None
    if dist is not None:
        gather, gathered = gather_legs(args, torch, dist, ctx, stream, step_dev, variants, S, g, ids, world, rank, barrier)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, log)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic",
            "config": {"workload": f"batch {S} synthetic 720p streams per GPU, {args.mode}-P picture step: fused MC+IDCT+add + YCbCr->RGBA "
                                   "(BASELINE configs[2]; stream-parallel over GPUs as configs[4])",
                       "streams_per_gpu": S, "macroblocks_per_step_per_gpu": int(len(mbs)), "coded_blocks_per_step_per_gpu": int(len(coeffs)),
                       "cache": "inputs larger than L2: %.0f MB of records + %.0f MB of reference frames per step vs 126 MB L2"
                                % ((h_mbs.numel() + h_coeffs.numel()) / 1e6, S * g.picture_bytes / 1e6),
                       "parallelism": f"stream-parallel x{world}, no data-path collective", "fused_ms": fused_ms,
                       "fused_frames_per_sec_per_gpu": S / (fused_ms * 1e-3)},
            "clocks": clocks, "parity": parity, "e2e": e2e, "e2e_int16": e2e_int16, "e2e_rgba_back": e2e_rgba, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (traffic or {}).get("dram_bytes_per_launch") if traffic else None,
                         "kernel": "plan_kernel + fused_tma_kernel (one mpegb200_video_decode_pictures_dev call)", "algorithmic_bytes_per_launch": int(alg_total),
                         "algorithmic_read_bytes_per_launch": int(alg_read), "read_only_frac": alg_read / (fused_ms * 1e-3) / 1e9 / peak,
                         "peak_source": peak_src, "launch_ms": fused_ms,
                         "dominant_kernel": {"kernel": "fused_tma_kernel alone (events between the pre-pass and the arithmetic kernel, second pass of the same steps)",
                                             "launch_ms": fused_only_ms, "plan_kernel_ms": plan_only_ms,
                                             "achieved": alg_total / (fused_only_ms * 1e-3) / 1e9,
                                             "frac": alg_total / (fused_only_ms * 1e-3) / 1e9 / peak}},
            "cpu_baseline": cpu,
        }
        line.update(extras)
        if gather:
            line["gather"] = gather
        if gathered:
            line["gathered"] = gathered
            line["value_gathered"] = gathered["value"]
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    ctx.close()


# ------------------------------------------------------------------------------------------------
# parity of the benchmarked configuration
# ------------------------------------------------------------------------------------------------
def oracle_frameset(ol, ctx, g, S):
    """An oracle FrameSet holding exactly what the GPU's frame buffers hold now (all three buffers of every stream)."""
    fs = ol.FrameSet(S, g.width, g.height)
    for s in range(S):
        for b in range(3):
            fs.whole(s, b)[:] = ctx.video_read_frame(s, b)
    return fs


def compare_planes(ctx, fs, g, S, ids, dst, h_planes, label, bad):
    ctx.video_read_pictures(ids, np.full(S, dst, np.uint8), h_planes.data_ptr(), g.picture_bytes)
    ctx.sync()
    got = h_planes.numpy().reshape(S, g.picture_bytes)
    for s in range(S):
        if not np.array_equal(got[s], fs.whole(s, dst)[:g.picture_bytes]):
            bad.append(f"{label}: stream {s} planes differ")
            if len(bad) > 8:
                break


def check_parity(args, ctx, g, S, refs, variants, mbs, coeffs, step_dev, decode_host, d_rgba, h_planes, ids, log):
    """Every stream of the batch: device-pointer path, RGBA, then the host-pointer (vlen and int16) paths continuing the same
    chain of P pictures, each against the CPU oracle executing the same records (mpeg_test.go:203-231 is the model: decode
    everything, compare everything)."""
    import oracle_lib as ol
    t0 = time.time()
    threads = host_threads()
    fs = ol.FrameSet(S, g.width, g.height)
    for s in range(S):
        fs.whole(s, 1)[:] = refs[s]
        fs.whole(s, 0)[:] = refs[(s + 1) % S]
        fs.whole(s, 2)[:] = refs[(s + 2) % S]
    bad, checks = [], []
    # 1. device-pointer path + RGBA (the step that `value` times)
    step_dev(0)
    p0, dst0 = variants[0]
    assert fs.exec_pictures(p0, mbs, coeffs, threads=threads) == 0
    compare_planes(ctx, fs, g, S, ids, dst0, h_planes, "device path", bad)
    checks.append("device-pointer decode")
    want = np.empty((S, g.height, g.width, 4), np.uint8)
    fs.rgba_batch(ids, np.full(S, dst0, np.uint8), want, threads=threads)
    got = d_rgba.cpu().numpy().reshape(S, g.height, g.width, 4)
    if not np.array_equal(got, want):
        bad.append("RGBA differs: streams " + str([int(s) for s in range(S) if not np.array_equal(got[s], want[s])][:8]))
    checks.append("rgba")
    del want, got
    # 2. host-pointer paths continue the chain: picture 2 through the vlen form, picture 3 through int16 blocks
    if not args.no_e2e:
        for k, form in ((1, "vlen"), (2, "int16")):
            dst = decode_host(k, form)
            pk, dk = variants[k]
            assert dk == dst and fs.exec_pictures(pk, mbs, coeffs, threads=threads) == 0
            compare_planes(ctx, fs, g, S, ids, dst, h_planes, f"host path ({form})", bad)
            checks.append(f"host-pointer decode ({form})")
    fs.close()
    ok = not bad
    log(f"parity of the benchmarked configuration: {'ok' if ok else 'MISMATCH ' + '; '.join(bad)} ({S} streams, {time.time() - t0:.1f} s)")
    return {"ok": ok, "checked_streams": S, "checks": checks, "against": "CPU oracle (oracle/, pinned by the reference's golden hashes) on the same records",
            "mismatches": bad}


# ------------------------------------------------------------------------------------------------
# sub-records
# ------------------------------------------------------------------------------------------------
def sustained_leg(args, torch, stream, step_dev, device_index, S, alg_total, peak):
    """The headline step back to back for --sustain-seconds (the default timed region lasts ~11 ms, a burst): ms per step,
    the decode call on a sample of the steps, SM clock and throttle reasons over the whole loop."""
    if args.sustain_seconds <= 0:
        return None
    sampler = ClockSampler(device_index)
    sampler.start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evs = []
    torch.cuda.synchronize()
    t_end = time.perf_counter() + args.sustain_seconds
    a.record(stream)
    k = 0
    while time.perf_counter() < t_end:
        for _ in range(32):
            if k % 16 == 0:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                evs.append(ev)
                step_dev(k, ev)
            else:
                step_dev(k)
            k += 1
        if k % 256 == 0:
            stream.synchronize()   # bound the queue depth
    b.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = a.elapsed_time(b)
    dec = [x.elapsed_time(y) for x, y in evs]
    half = len(dec) // 2
    dec_ms = float(np.mean(dec[half:])) if dec else None   # second half: clocks have settled
    return {"seconds": ms * 1e-3, "steps": k, "ms_per_step": ms / k, "value": S * k / (ms * 1e-3), "decode_call_ms": dec_ms,
            "roofline_frac": (alg_total / (dec_ms * 1e-3) / 1e9 / peak) if dec_ms else None, "clocks": clocks,
            "note": "decode_call_ms = mean over the second half of the loop, every 16th step bracketed by events"}


def picture_steps(torch, ctx, stream, g, S, peak, log):
    """The other picture steps of BASELINE configs[2], 256 streams each: parity against the oracle, then the decode
    call timed like the headline step; fraction of the HBM roofline on the step's OWN algorithmic bytes."""
    import oracle_lib as ol
    import workload as wl
    out = []
    threads = host_threads()
    ids = np.arange(S, dtype=np.int32)
    h_planes = torch.empty(S * g.picture_bytes, dtype=torch.uint8, pin_memory=True)
    table = [("natural-P", wl.PIC_P, "natural", 32), ("natural-B", wl.PIC_B, "natural", 32), ("I-only", wl.PIC_I, "natural", 32),
             ("dense-P wide vectors (+-64 px)", wl.PIC_P, "dense", 128)]
    for name, ptype, mode, mv_range in table:
        t0 = time.time()
        _, _, pics, mbs, coeffs = build_batch(S, 0, mode, None, ptype, mv_range, seed_offset=5000, with_refs=False)
        alg, _ = wl.algorithmic_bytes(mbs, len(coeffs))
        ctx.video_validate(pics, mbs, len(coeffs))
        d = [torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).cuda() for a in (pics, mbs, coeffs)]

        def call():
            with torch.cuda.stream(stream):
                ctx.video_decode_pictures_dev(len(pics), d[0].data_ptr(), len(mbs), d[1].data_ptr(), len(coeffs), d[2].data_ptr())

        # parity: the oracle starts from the GPU's current frame buffers and executes the same records
        ctx.sync()
        fs = oracle_frameset(ol, ctx, g, S)
        call()
        assert fs.exec_pictures(pics, mbs, coeffs, threads=threads) == 0
        bad = []
        compare_planes(ctx, fs, g, S, ids, 0, h_planes, name, bad)
        fs.close()
        # timing: per-call events; every call reads its 400-723 MB of records + 354 MB of frames again (larger than the L2)
        n_t = 10
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_t)]
        for k in range(3 + n_t):
            if k >= 3:
                evs[k - 3][0].record(stream)
            call()
            if k >= 3:
                evs[k - 3][1].record(stream)
        torch.cuda.synchronize()
        ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
        rec = {"step": name, "streams": S, "macroblocks": int(len(mbs)), "coded_blocks": int(len(coeffs)),
               "predicted_frac": round(float(((mbs["flags"] & wl.MB_PREDICT) != 0).mean()), 3), "algorithmic_bytes": int(alg),
               "decode_call_ms": ms, "frames_per_sec": S / (ms * 1e-3), "gbps": alg / ms / 1e6, "roofline_frac": alg / ms / 1e6 / peak,
               "parity_ok": not bad, "mismatches": bad}
        log(f"step {name}: {ms:.4f} ms, {rec['roofline_frac']:.3f} of the HBM roofline on its bytes, parity {'ok' if not bad else 'MISMATCH'} "
            f"({time.time() - t0:.1f} s)")
        out.append(rec)
        del d
        if bad:
            fail_parity({"metric": METRIC, "error": f"picture step {name}: GPU output differs from the CPU oracle", "mismatches": bad})
    return out


def audio_leg(torch, device_index, peak, log):
    """BASELINE configs[3]: 1024 MP2 streams x 8 frames per launch (idct36 + synthesis window + scaling), both window modes."""
    import mpeg_b200
    import oracle_lib as ol
    import workload as wl
    n_s, n_f = AUDIO_STREAMS, AUDIO_FRAMES
    rng = wl.stream_rng(4, 0)
    samples = wl.audio_samples(rng, n_s * n_f)
    ids = np.arange(n_s, dtype=np.int32)
    threads = host_threads()
    stream = torch.cuda.Stream()
    out = {"config": f"{n_s} streams x {n_f} frames per launch (BASELINE configs[3]), samples and output resident in HBM",
           "algorithmic_bytes_per_frame": AUDIO_BYTES_PER_FRAME, "modes": {}}
    with mpeg_b200.Context(device=device_index, max_streams=n_s) as actx:
        actx.set_stream(stream.cuda_stream)
        d_samples = torch.from_numpy(samples.reshape(-1)).cuda()
        d_out = torch.empty(n_s * n_f * 2304, dtype=torch.float32, device="cuda")
        for mode, flag in (("unfused (the reference's Go / SSE window, hash 0xf1b76cdf8e6cdea5)", 0),
                           ("fused (the reference's AVX2 / NEON window, hash 0x50f3ab75f5fb0fb5)", mpeg_b200.AUDIO_WINDOW_FMA)):
            for s in range(n_s):
                actx.audio_open(s)
            # parity on the first launch (fresh states), every sample of every stream
            with torch.cuda.stream(stream):
                actx.audio_synth_dev(ids, n_f, d_samples.data_ptr(), mpeg_b200.AUDIO_F32N | flag, d_out.data_ptr())
            stream.synchronize()
            got = d_out.cpu().numpy().reshape(n_s, n_f, 2304)
            t_c = time.perf_counter()
            want = ol.synth_batch(ol.synth_states(n_s), n_s, n_f, samples, 0, fma=bool(flag), threads=threads)
            cpu_s = time.perf_counter() - t_c
            ok = bool(np.array_equal(got.view(np.uint32), want.view(np.uint32)))
            n_t = 20
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_t)]
            with torch.cuda.stream(stream):
                for k in range(3 + n_t):
                    if k >= 3:
                        evs[k - 3][0].record(stream)
                    actx.audio_synth_dev(ids, n_f, d_samples.data_ptr(), mpeg_b200.AUDIO_F32N | flag, d_out.data_ptr())
                    if k >= 3:
                        evs[k - 3][1].record(stream)
            torch.cuda.synchronize()
            ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
            alg = n_s * n_f * AUDIO_BYTES_PER_FRAME
            key = mode.split(" ")[0]
            out["modes"][key] = {
                "mode": mode, "ms_per_launch": ms, "frames_per_sec": n_s * n_f / (ms * 1e-3), "gbps": alg / ms / 1e6,
                "roofline_frac": alg / ms / 1e6 / peak, "parity_ok": ok,
                "cpu_frames_per_sec": n_s * n_f / cpu_s, "cpu_threads": threads}
            log(f"audio {key}: {ms:.4f} ms per launch, {alg / ms / 1e6 / peak:.3f} of the HBM roofline, parity {'ok' if ok else 'MISMATCH'}")
            for s in range(n_s):
                actx.audio_close(s)
            if not ok:
                fail_parity({"metric": METRIC, "error": f"audio ({mode}): GPU samples differ from the CPU oracle"})
    return out


def bitstream_leg(args, device_index, log, paths=("host_parser", "device_vlc", "device_vlc_resident", "device_vlc_resident_to_host"), barrier=None,
                  max_over_ranks=None, world=1):
    """Bitstream in, frames out: S natural 720p streams (two distinct ones written by tests/mpeg1_writer.py, I + P pictures)
    through mpeg_b200.VideoBatch with the host parser and with the device-side slice parser.  Frames stay in HBM (the frames of four
    streams are read back every step, outside the rate, and hashed against the oracle's decode of the same bitstream), except in the
    `_to_host` path, which is the end-to-end form: every returned frame of every stream travels to pinned host memory inside the
    timed region (its parity is taken in a second, untimed pass that waits for every copy)."""
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor
    import mpeg_b200
    import oracle_lib as ol
    sys.path.insert(0, str(ROOT / "tools"))
    import bench_bitstream as bb
    S = args.streams
    cached = sorted((ROOT / "tools" / "_build" / "streams").glob("natural_720p_seed72[01]_40pictures.m1v"))
    if len(cached) == 2:
        distinct = [f.read_bytes() for f in cached]
    else:
        distinct = bb.make_streams(2, 16, "natural", log)   # no cached streams on this box: two worker processes write them (about ten seconds)
    n_pictures = sum(1 for i in range(len(distinct[0]) - 3) if distinct[0][i:i + 4] == b"\x00\x00\x01\x00")
    streams = [distinct[i % 2] for i in range(S)]
    threads = host_threads()

    def oracle_decode(d):   # the oracle's full decoder on the bitstream (parse + reconstruct), hash over all returned frames
        o, h, n = ol.VideoOracle(d), ol.FNV_OFFSET, 0
        while (f := o.decode()) is not None:
            for which in ("y", "cb", "cr"):
                h = ol.fnv(h, f.plane(which))
            n += 1
        return h, n

    want = [oracle_decode(d)[0] for d in distinct]
    out = {"config": f"{S} natural 720p streams of {n_pictures} pictures ({len(distinct[0]) / n_pictures / 1e3:.1f} KB per picture)",
           "host_threads": threads, "paths": {}}
    if world == 1:
        # the CPU restatement of the reference's decoder on the same bitstreams: one stream per host thread, bit-serial VLC walk like the Go code
        def cpu_decode(d):
            o, n = ol.VideoOracle(d), 0
            while o.decode() is not None:
                n += 1
            return n
        t0 = time.perf_counter()
        with ThreadPoolExecutor(threads) as ex:   # ctypes releases the GIL inside the oracle
            n_cpu = sum(ex.map(cpu_decode, [distinct[i % 2] for i in range(threads)]))
        cpu_s = time.perf_counter() - t0
        out["cpu_decoder"] = {"frames_per_sec": n_cpu / cpu_s, "threads": threads, "kind": "port",
                              "what": "oracle/ full decoder (parse + reconstruct) on the same bitstreams, one stream per thread, planes in host memory"}
        log(f"bitstream cpu decoder: {n_cpu / cpu_s:.0f} frames/s on {threads} threads")
    check = [0, 1, S - 2, S - 1] if S >= 4 else list(range(S))
    L = mpeg_b200._lib.load()
    all_paths = {"host_parser": {}, "device_vlc": {"device_vlc": True}, "device_vlc_resident": {"device_vlc": True, "resident": True},
                 "device_vlc_resident_to_host": {"device_vlc": True, "resident": True}}
    if world > 1:
        threads = max(1, threads // world)   # the ranks share the host's cores
        out["host_threads"] = threads
    for name in paths:
        kw = all_paths[name]
        to_host = name.endswith("_to_host")
        for timed_pass in ((True, False) if to_host else (True,)):
            with mpeg_b200.Context(device=device_index, max_streams=S) as c:
                c.set_kernel_timing(name != "host_parser")
                vb = mpeg_b200.VideoBatch(c, streams, threads=threads, validate=False, **kw)
                geo = c.video_geometry(0)
                pic_bytes = geo[0] * geo[1] + 2 * geo[2] * geo[3]
                host = np.empty((len(check), pic_bytes), np.uint8)
                if to_host:
                    import torch
                    h_all = [torch.empty((S, pic_bytes), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
                    all_ids = np.arange(S, dtype=np.int32)
                hashes = [ol.FNV_OFFSET] * len(check)
                frames, parse_ms, ms, steps = 0, [], C.c_float(), 0
                t_hash = 0.0
                c.sync()
                if barrier:
                    barrier()
                t0 = time.perf_counter()
                while True:
                    has, buf, _ = vb.step()
                    if not has.any():
                        break
                    frames += int(has.sum())
                    if name != "host_parser" and L.mpegb200_video_bitstream_parse_ms(c.h, C.byref(ms)) == 0:
                        parse_ms.append(ms.value)
                    if to_host:
                        live = np.nonzero(has)[0]
                        dst = h_all[steps & 1]
                        c.video_read_pictures(all_ids[live], buf[live], dst.data_ptr(), pic_bytes)   # asynchronous, on the read-back stream
                        if not timed_pass:   # parity pass: wait for the copy and hash what arrived on the host
                            c.sync()
                            for k, i in enumerate(check):
                                if has[i]:
                                    hashes[k] = ol.fnv(hashes[k], dst[int(np.searchsorted(live, i))].numpy())
                    else:
                        th = time.perf_counter()   # parity read-back of a few streams: outside the rate
                        live = [k for k, i in enumerate(check) if has[i]]
                        if live:
                            c.video_read_pictures(np.array([check[k] for k in live]), buf[[check[k] for k in live]], host.ctypes.data, pic_bytes)
                            c.sync()
                            for j, k in enumerate(live):
                                hashes[k] = ol.fnv(hashes[k], host[j])
                        t_hash += time.perf_counter() - th
                    steps += 1
                c.sync()   # waits for the read-back stream too
                dt = time.perf_counter() - t0 - t_hash
                if max_over_ranks:
                    dt = max_over_ranks(dt * 1e3) * 1e-3   # the slowest rank
                    frames *= world
                if to_host and timed_pass:
                    timed = (frames, dt, list(parse_ms), vb.flagged, vb.host_steps, (vb.t_scan, vb.t_submit, vb.t_wait))
                    vb.close()
                    continue
                if to_host:
                    frames, dt, parse_ms, flagged, host_steps, t_in = timed
                else:
                    flagged, host_steps, t_in = vb.flagged, vb.host_steps, (vb.t_scan, vb.t_submit, vb.t_wait)
                ok = all(hashes[k] == want[check[k] % 2] for k in range(len(check)))
                rec = {"frames_per_sec": frames / dt, "frames": frames, "seconds": dt, "parity_ok": bool(ok), "checked_streams": check}
                if name != "host_parser":
                    steady = sorted(parse_ms[1:] or parse_ms)
                    rec.update({"flagged_pictures": flagged, "host_steps": host_steps,
                                "parse_kernel_ms_per_wave": steady[len(steady) // 2] if steady else None,
                                "parse_kernel_pictures_per_sec": S / (steady[len(steady) // 2] * 1e-3) if steady else None,
                                "seconds_in": {"host_scan": t_in[0], "submit": t_in[1], "waiting_for_flags": t_in[2]}})
                if to_host:
                    rec["d2h_bytes_per_step"] = int(S * pic_bytes)
                    rec["note"] = ("end to end from the bitstream: streams resident in HBM, every returned frame copied to pinned host memory inside the timed "
                                   "region (double-buffered, asynchronous); parity from a second pass that waits for every copy")
                out["paths"][name] = rec
                log(f"bitstream {name}: {frames / dt:.0f} frames/s, parity {'ok' if ok else 'MISMATCH'}")
                vb.close()
                if not ok:
                    fail_parity({"metric": METRIC, "error": f"bitstream ({name}): frames differ from the oracle's decode of the bitstream"})
    if "host_parser" in out["paths"]:
        out["device_vlc_speedup"] = out["paths"]["device_vlc"]["frames_per_sec"] / out["paths"]["host_parser"]["frames_per_sec"]
        out["device_vlc_resident_speedup"] = out["paths"]["device_vlc_resident"]["frames_per_sec"] / out["paths"]["host_parser"]["frames_per_sec"]
    if "cpu_decoder" in out and "device_vlc_resident_to_host" in out["paths"]:
        out["to_host_vs_cpu_decoder"] = out["paths"]["device_vlc_resident_to_host"]["frames_per_sec"] / out["cpu_decoder"]["frames_per_sec"]
    if world > 1:
        out["config"] += f"; {world} ranks, whole-job frames/s over the slowest rank's time"
    if "device_vlc_resident" in out["paths"]:
        out["paths"]["device_vlc_resident"]["note"] = ("streams uploaded to HBM and their start codes indexed on the device when the batch is created (outside the rate, "
                                                        "like the demux); per step the host reads headers and builds slice tables, no compressed byte crosses PCIe")
    return out


def transfer_probe(torch):
    """What the end-to-end leg has to live with on this box: this rank's H2D and D2H rates (pinned, 256 MiB, alone and both
    directions at once) and the host's memcpy bandwidth (one thread, and all of this rank's threads)."""
    n = 256 << 20
    h_a = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_b = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_a.zero_()
    h_b.zero_()
    d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_b = torch.zeros(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(up, down):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4):
            if up:
                with torch.cuda.stream(s1):
                    d_a.copy_(h_a, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    h_b.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        return 4 * n / (time.perf_counter() - t0) / 1e9

    run(True, True)
    res = {"h2d_gbs": run(True, False), "d2h_gbs": run(False, True)}
    res["bidirectional_gbs_each"] = run(True, True)
    # host memory: numpy copies of 256 MiB (the GIL is released), one thread and all threads of this rank
    from concurrent.futures import ThreadPoolExecutor
    src = np.ones(n, np.uint8)
    dst = np.empty(n, np.uint8)
    np.copyto(dst, src)
    t0 = time.perf_counter()
    for _ in range(3):
        np.copyto(dst, src)
    res["host_memcpy_gbs_1_thread"] = 3 * n / (time.perf_counter() - t0) / 1e9
    th = max(1, min(host_threads(), 32))
    bounds = np.linspace(0, n, th + 1).astype(np.int64) // 4096 * 4096
    bounds[-1] = n

    def part(i):
        np.copyto(dst[bounds[i]:bounds[i + 1]], src[bounds[i]:bounds[i + 1]])
    with ThreadPoolExecutor(th) as ex:
        list(ex.map(part, range(th)))
        t0 = time.perf_counter()
        for _ in range(3):
            list(ex.map(part, range(th)))
        res["host_memcpy_gbs_all_threads"] = 3 * n / (time.perf_counter() - t0) / 1e9
    res["host_threads"] = th
    try:
        res["numa_nodes"] = len([p for p in os.listdir("/sys/devices/system/node") if p.startswith("node")])
    except Exception:
        res["numa_nodes"] = None
    return res


def gather_legs(args, torch, dist, ctx, stream, step_dev, variants, S, g, ids, world, rank, barrier):
    """(1) the NCCL gather of one step's planes alone; (2) `gathered`: K steps with every step's decoded planes travelling to rank 0
    over NVLink INSIDE the timed region, double-buffered so that the transfer of step k overlaps the decode of step k + 1."""
    from mpeg_b200.sharding import gather_frames
    send = [torch.empty((S, g.picture_bytes), dtype=torch.uint8, device="cuda") for _ in range(2)]
    recv = [torch.empty((world, S, g.picture_bytes), dtype=torch.uint8, device="cuda") for _ in range(2)] if rank == 0 else None
    with torch.cuda.stream(stream):
        ctx.video_read_pictures(ids, np.full(S, variants[0][1], np.uint8), send[0].data_ptr(), g.picture_bytes, device=True)
    stream.synchronize()
    gather_frames(send[0], dst=0)  # warm-up
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    gather_frames(send[0], dst=0)
    b.record()
    barrier()
    t = torch.tensor([a.elapsed_time(b)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gather = {"ms": float(t.item()), "bytes_per_rank": int(send[0].numel()), "note": "NCCL gather of one step's decoded "
              "Y/Cb/Cr planes to rank 0, alone (root ingress bound)"}

    pending = [None, None]

    def wait_slot(slot):
        if pending[slot] is not None:
            for w in pending[slot]:
                w.wait()                        # makes the current stream wait for the transfer
            pending[slot] = None

    def step_gathered(k):
        slot = k & 1
        with torch.cuda.stream(stream):
            wait_slot(slot)                     # the slot's previous transfer must have left before it is refilled
        step_dev(k)
        dst = variants[k % 3][1]
        with torch.cuda.stream(stream):
            ctx.video_read_pictures(ids, np.full(S, dst, np.uint8), send[slot].data_ptr(), g.picture_bytes, device=True)
            # issued from the decode stream: NCCL's stream waits for the pack above, the decode of step k + 1 does not wait for NCCL
            if args.gather_mode == "gather":
                works = [dist.gather(send[slot], list(recv[slot].unbind(0)) if rank == 0 else None, dst=0, async_op=True)]
            elif rank == 0:
                recv[slot][0].copy_(send[slot], non_blocking=True)
                works = dist.batch_isend_irecv([dist.P2POp(dist.irecv, recv[slot][r], r) for r in range(1, world)])
            else:
                works = dist.batch_isend_irecv([dist.P2POp(dist.isend, send[slot], 0)])
        pending[slot] = works

    for k in range(3):
        step_gathered(k)
    with torch.cuda.stream(stream):
        wait_slot(0)
        wait_slot(1)
    steps = max(3, args.steps)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    a.record(stream)
    for k in range(steps):
        step_gathered(k)
    with torch.cuda.stream(stream):
        wait_slot(0)
        wait_slot(1)
    b.record(stream)
    barrier()
    t = torch.tensor([a.elapsed_time(b)], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gms = float(t.item())
    gathered = {"value": S * world * steps / (gms * 1e-3), "unit": UNIT, "ms_per_step": gms / steps, "steps": steps, "mode": args.gather_mode,
                "bytes_into_rank0_per_step": int((world - 1) * send[0].numel()),
                "note": "decode + RGBA + pack + transfer of the planes of all ranks to rank 0 inside the timed region; transfers double-buffered "
                        "(p2p: one ncclRecv per peer in one group; gather: ncclGather), max over ranks"}
    return gather, gathered


if __name__ == "__main__":
    main()
