#!/usr/bin/env python3
"""bench.py -- MPEG-1 720p frames/sec (batched) on N B200s, with the roofline of the fused kernel.

One "step" = one pass of the hot path over one batch of synthetic pre-parsed input: per GPU 256
independent 720p streams each decode one dense-P picture (fused motion compensation + 8x8 IDCT +
residual add, one launch) and convert the decoded frame YCbCr->RGBA (one launch).  This is
BASELINE.json configs[2] ("batch 256 synthetic 720p streams on 1 B200"); with --gpus N every rank
holds its own 256 streams (configs[4] layout: stream-parallel, no data-path collective, weak scaling).

  value     whole-job frames/s with the packed records already resident in HBM (device-timed)
  e2e       the same step through the host-pointer C-ABI: pinned host records -> H2D -> kernels ->
            D2H of the decoded Y/Cb/Cr planes into pinned host memory, all inside the timed region.  The records
            are "pre-parsed block batches" in the transfer form the entry point takes (--e2e-form: variable-width
            groups by default); producing that form from int16 blocks is host work done once before the loop, its
            duration is reported as e2e.host_pack_ms_once
            roofline.dominant_kernel = the arithmetic kernel alone (library events between pre-pass and kernel)
  roofline  fused MC+IDCT+add kernel: algorithmic bytes (1552 B per dense-P macroblock)
            / mean launch duration measured with CUDA events on the launching stream, against the
            measured HBM peak of MEASURED_PEAKS.json
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
    ap.add_argument("--e2e-form", default="vlen", choices=["vlen", "packed12", "int16"], help="coefficient transfer form of the end-to-end leg")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="target CPU work for cpu_baseline")
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
        self.samples, self.reason_bits, self._stop = [], 0, threading.Event()

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
                    "reasons": reasons, "samples": len(self.samples), "source": "nvml"}
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


def build_batch(streams, first_stream_id, mode, log=None):
    """Packed records of one picture step for `streams` 720p streams (one P picture each)."""
    import workload as wl
    g = wl.HD720
    per, refs = [], []
    for s in range(streams):
        rng = wl.stream_rng(CONFIG_ID, first_stream_id + s)
        refs.append(wl.random_reference_frame(rng, g))
        per.append(wl.make_picture(rng, g, wl.PIC_P, mode))
        if log and (s + 1) % 64 == 0:
            log(f"generated {s + 1}/{streams} streams")
    pics, mbs, coeffs = wl.batch_pictures(per, list(range(streams)), wl.PIC_P, [(0, 1, 2)] * streams)
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
def cpu_step(ol, fs, pics, mbs, coeffs, threads, with_rgba=True):
    t0 = time.perf_counter()
    rc = fs.exec_pictures(pics, mbs, coeffs, threads=threads)
    assert rc == 0
    if with_rgba:
        import ctypes as C
        from concurrent.futures import ThreadPoolExecutor
        out = [np.empty((fs.height, fs.width, 4), np.uint8) for _ in range(len(pics))]

        def one(i):
            ol.lib().orc_rgba(C.byref(fs.frame(int(pics["stream"][i]), int(pics["dst_buf"][i]))), out[i].ctypes.data)
        with ThreadPoolExecutor(max_workers=threads) as ex:  # ctypes releases the GIL
            list(ex.map(one, range(len(pics))))
    return time.perf_counter() - t0


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_setup(n_pictures, mode):
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib as ol
    g, refs, pics, mbs, coeffs = build_batch(n_pictures, 0, mode)
    fs = ol.FrameSet(n_pictures, g.width, g.height)
    for s in range(n_pictures):
        fs.whole(s, 1)[:] = refs[s]
    return ol, fs, pics, mbs, coeffs


def best_thread_count(ol, fs, pics, mbs, coeffs, n):
    """All hardware threads are not always the fastest (SMT siblings share the integer units): give the CPU
    arm whichever of n and n/2 threads is faster on one trial step."""
    cands = [n] if n < 4 else [n, n // 2]
    timing = {t: min(cpu_step(ol, fs, pics, mbs, coeffs, t) for _ in range(2)) for t in cands}
    return min(timing, key=timing.get)


def run_reference(args):
    """--impl reference: rank 0 only; each step = one picture per host thread of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = host_threads()
    ol, fs, pics, mbs, coeffs = cpu_setup(n, args.mode)
    threads = best_thread_count(ol, fs, pics, mbs, coeffs, n)
    for _ in range(args.warmup):
        cpu_step(ol, fs, pics, mbs, coeffs, threads)
    t = sum(cpu_step(ol, fs, pics, mbs, coeffs, threads) for _ in range(args.steps))
    fps = len(pics) * args.steps / t
    sample = f"{len(pics)} dense-P 720p pictures per step (one per host thread) + RGBA, {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": f"720p {args.mode}-P fused MC+IDCT+add + YCbCr->RGBA, CPU restatement of the reference Go path "
                               "(Go toolchain absent), sample of the 256-stream batch", "pictures_per_step": len(pics)},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample + f" ({threads} of {n} hardware threads: the faster of n and n/2)"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def cpu_baseline(args, log):
    n = host_threads()
    ol, fs, pics, mbs, coeffs = cpu_setup(n, args.mode)
    threads = best_thread_count(ol, fs, pics, mbs, coeffs, n)
    t1 = cpu_step(ol, fs, pics, mbs, coeffs, threads)  # warm-up / calibration
    rounds = int(max(2, min(200, args.cpu_seconds / max(t1, 1e-3))))
    t = sum(cpu_step(ol, fs, pics, mbs, coeffs, threads) for _ in range(rounds))
    fps = len(pics) * rounds / t
    log(f"cpu baseline: {fps:.1f} frames/s on {threads} threads ({rounds} rounds, {t:.1f} s)")
    return {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{rounds} rounds x {len(pics)} dense-P 720p pictures (fused-kernel-equivalent replay + RGBA), "
                      f"{t:.1f} s wall, OpenMP over pictures"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
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
    # transfer form of the coefficients on the end-to-end path (PCIe bound): variable-width groups (about 49 B per dense
    # block), the fixed 12-bit form (96 B) or the plain int16 blocks (128 B)
    h_packed = h_vlen = None
    pack_ms = None
    if not args.no_e2e:
        try:
            if args.e2e_form == "vlen":
                t_pack = time.perf_counter()
                packed_arrays = ctx.pack_coeffs_vlen(coeffs)
                pack_ms = 1e3 * (time.perf_counter() - t_pack)   # host work outside the timed region, reported next to e2e
                h_vlen = tuple(pinned(a) for a in packed_arrays)
            elif args.e2e_form == "packed12":
                h_packed = pinned(ctx.pack_coeffs12(coeffs))
        except mpeg_b200.MpegB200Error:
            h_packed = h_vlen = None
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

    def step_e2e(k):
        p, dst = variants[k % 3]
        L = ctx.L
        if h_vlen is not None:
            rc = L.mpegb200_video_decode_pictures_vlen(ctx.h, len(p), h_pics[k % 3].data_ptr(), len(mbs), h_mbs.data_ptr(), len(coeffs),
                                                       h_vlen[0].data_ptr(), h_vlen[1].data_ptr(), h_vlen[2].data_ptr(), h_vlen[2].numel())
        elif h_packed is not None:
            rc = L.mpegb200_video_decode_pictures_packed(ctx.h, len(p), h_pics[k % 3].data_ptr(), len(mbs), h_mbs.data_ptr(), len(coeffs), h_packed.data_ptr())
        else:
            rc = L.mpegb200_video_decode_pictures(ctx.h, len(p), h_pics[k % 3].data_ptr(), len(mbs), h_mbs.data_ptr(), len(coeffs), h_coeffs.data_ptr())
        ctx._ck(rc)
        ctx.video_rgba_batch_dev(ids, np.full(S, dst, np.uint8), d_rgba.data_ptr(), rgba_stride)
        ctx.video_read_pictures(ids, np.full(S, dst, np.uint8), h_planes.data_ptr(), g.picture_bytes)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

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
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, kms

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
    peak, peak_src = measured_peak()
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

    # ---- end-to-end run (host buffers, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        for k in range(max(1, args.warmup)):
            step_e2e(k)
        e_steps = max(3, args.steps)  # same K as the device-resident run: the fill and drain of the three-stream pipeline are part of it
        ems, _ = timed(step_e2e, e_steps)
        e2e = {"value": S * world * e_steps / (ems * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(h_mbs.numel() + h_pics[0].numel() + (sum(t.numel() for t in h_vlen) if h_vlen is not None else
                                                                              (h_packed if h_packed is not None else h_coeffs).numel())),
               "d2h_bytes_per_step": int(h_planes.numel()), "steps": e_steps, "ms_per_step": ems / e_steps,
               "host_pack_ms_once": pack_ms, "host_pack_threads": min(64, os.cpu_count() or 1) if pack_ms is not None else None,
               "path": ("mpegb200_video_decode_pictures_vlen (pinned host records, variable-width coefficient transfer form)" if h_vlen is not None
                        else "mpegb200_video_decode_pictures_packed (pinned host records, 12-bit coefficient transfer form)" if h_packed is not None
                        else "mpegb200_video_decode_pictures (pinned host records)") + " + rgba_batch_dev + read_pictures_host (pinned)"}

    # ---- NCCL gather of the decoded frames (the only collective of the path), timed on its own
    gather = None
    if dist is not None:
        from mpeg_b200.sharding import gather_frames
        send = torch.empty((S, g.picture_bytes), dtype=torch.uint8, device="cuda")
        with torch.cuda.stream(stream):
            ctx.video_read_pictures(ids, np.full(S, variants[0][1], np.uint8), send.data_ptr(), g.picture_bytes, device=True)
        stream.synchronize()
        gather_frames(send, dst=0)  # warm-up
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gather_frames(send, dst=0)
        b.record()
        barrier()
        t = torch.tensor([a.elapsed_time(b)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gather = {"ms": float(t.item()), "bytes_per_rank": int(send.numel()), "note": "NCCL gather of one step's decoded "
                  "Y/Cb/Cr planes to rank 0, outside the timed region (SURVEY 8e: root ingress bound)"}

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
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
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
        if gather:
            line["gather"] = gather
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
