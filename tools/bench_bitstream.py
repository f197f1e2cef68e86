#!/usr/bin/env python3
"""Bitstream -> frames at 720p: does the host parser starve the GPU?

Writes a few synthetic 720p MPEG-1 elementary streams with tests/mpeg1_writer.py (shaped like BASELINE configs[2]: dense P
pictures -- every macroblock predicted, six coded blocks of 64 coefficients -- or "natural" ones), replicates them to N streams,
and runs the lock-step batch decoder (mpegb200_video_batch_*, parsers emitting the variable-width coefficient form):

  host     the parse + wave merge alone (no GPU needed): pictures/s, MB/s of bitstream, per host thread
  gpu      the same with every wave going through mpegb200_video_decode_pictures_vlen (needs a B200): frames/s from
           compressed bitstream to decoded frames in HBM

usage: bench_bitstream.py [--streams 256] [--threads T] [--mode dense|natural] [--pictures 4] [--distinct 4] [--gpu]
Prints one JSON line."""
import argparse
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def write_one(d, pictures, mode):
    """Stream number d: an I picture and pictures - 1 P pictures of 1280x720 (a top-level function: it also runs in worker processes)."""
    sys.path.insert(0, str(ROOT / "tests"))
    import mpeg1_writer as mw
    rng = np.random.default_rng(720 + d)
    w = mw.StreamWriter(1280, 720, quantizer_scale=8, f_code=2)
    w.picture(mw.PIC_I, mw.random_picture(rng, w.mb_w, w.mb_h, mw.PIC_I, mode, 32, 1280, 720))
    for _ in range(pictures - 1):
        w.picture(mw.PIC_P, mw.random_picture(rng, w.mb_w, w.mb_h, mw.PIC_P, mode, 32, 1280, 720))
    return w.tobytes()


def make_streams(distinct, pictures, mode, log):
    """Synthetic 720p streams (I then P pictures); cached under tools/_build/streams/ (git-ignored, travels to the GPU box) because
    the Python writer needs about half a second per picture.  Missing streams are written in parallel worker processes."""
    cache = ROOT / "tools" / "_build" / "streams"
    cache.mkdir(parents=True, exist_ok=True)
    files = [cache / f"{mode}_720p_seed{720 + d}_{pictures}pictures.m1v" for d in range(distinct)]
    missing = [d for d in range(distinct) if not files[d].exists()]
    if missing:   # one writer process per missing stream (plain subprocesses: nothing of the caller's state, CUDA or otherwise, is inherited)
        import subprocess
        t0 = time.time()
        procs = [subprocess.Popen([sys.executable, str(Path(__file__).resolve()), "--write-one", str(d), "--pictures", str(pictures), "--mode", mode,
                                   "--out", str(files[d])]) for d in missing]
        if any(p.wait() != 0 for p in procs):
            raise RuntimeError("writing the synthetic streams failed")
        log(f"{len(missing)} stream(s) of {pictures} pictures written in {time.time() - t0:.1f} s")
    out = [f.read_bytes() for f in files]
    for d, data in enumerate(out):
        log(f"stream {d}: {files[d].name}, {len(data) / 1e6:.2f} MB")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=256)
    ap.add_argument("--threads", type=int, default=0, help="host threads of the batch parser (default: all)")
    ap.add_argument("--mode", default="dense", choices=["dense", "natural"])
    ap.add_argument("--pictures", type=int, default=5)
    ap.add_argument("--distinct", type=int, default=2)
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--write-one", type=int, default=None, help="(internal) write stream number N to --out and exit")
    ap.add_argument("--out", default=None)
    ap.add_argument("--device-vlc", action="store_true", help="with --gpu: also run the slice-parallel VLC stage on the device")
    ap.add_argument("--resident", action="store_true", help="with --device-vlc: also with the streams resident in HBM and indexed there")
    args = ap.parse_args()
    if args.write_one is not None:
        tmp = Path(f"{args.out}.{os.getpid()}.tmp")   # ranks of one job may write the same stream at the same time: rename is atomic
        tmp.write_bytes(write_one(args.write_one, args.pictures, args.mode))
        os.replace(tmp, args.out)
        return
    threads = args.threads or len(os.sched_getaffinity(0))

    def log(m):
        print(f"[bitstream] {m}", file=sys.stderr, flush=True)

    distinct = make_streams(args.distinct, args.pictures, args.mode, log)
    if args.streams == 0:
        return
    streams = [distinct[i % len(distinct)] for i in range(args.streams)]
    total_bytes = sum(len(s) for s in streams)
    from mpeg_b200 import _lib
    from mpeg_b200.batch import BatchStep
    L = _lib.load()

    # ---- host side alone
    def host_run(n_threads):
        h = L.mpegb200_video_batch_new(len(streams), n_threads, None, None)
        L.mpegb200_video_batch_set_vlen(h, 1)
        for i, d in enumerate(streams):
            assert L.mpegb200_video_batch_set_stream(h, i, d, len(d)) == 0
        st = BatchStep()
        pics = blocks = payload = 0
        dt, k = 0.0, 0
        while True:
            t0 = time.perf_counter()
            assert L.mpegb200_video_batch_next(h, C.byref(st)) == 0
            t1 = time.perf_counter()
            got = int(np.ctypeslib.as_array(st.has_frame, shape=(len(streams),)).sum())
            if got == 0:
                break
            k += 1
            if k == 1:
                continue   # the first step grows every parser's arrays (allocation, page faults): steady state starts behind it
            dt += t1 - t0
            for w in range(st.n_waves):
                pics += st.waves[w].n_pictures
                blocks += st.waves[w].n_blocks
                payload += st.waves[w].vlen_payload_bytes + 4 * st.waves[w].n_blocks
        L.mpegb200_video_batch_free(h)
        return dt, pics, blocks, payload

    dt1, pics1, _, _ = host_run(1) if args.streams <= 32 else (None, 0, 0, 0)
    dtn, pics, blocks, payload = host_run(threads)
    res = {"what": "bitstream -> records (host parser, vlen form) and -> frames (GPU)", "mode": args.mode, "streams": args.streams,
           "pictures_per_stream": args.pictures, "bitstream_bytes_per_picture": total_bytes / (args.streams * args.pictures),
           "coded_blocks_per_picture": blocks / max(pics, 1), "record_bytes_per_picture_vlen": (payload + 16 * 3600 * pics) / max(pics, 1),
           "host": {"threads": threads, "pictures_per_sec": pics / dtn, "bitstream_MB_per_sec": pics * total_bytes / (args.streams * args.pictures) / dtn / 1e6,
                    "timed": "all steps but the first (which allocates)",
                    "pictures_per_sec_per_thread": pics / dtn / threads,
                    "single_thread_pictures_per_sec": (pics1 / dt1) if dt1 else None}}
    fused_fps = 814000.0   # fused MC+IDCT+add kernel alone, frames/s per GPU (bench.py, BENCH line)
    res["host"]["threads_to_feed_one_gpu_decode_kernel"] = fused_fps / (pics / dtn / threads)

    if args.gpu:
        import mpeg_b200
        ctx = mpeg_b200.Context(0, args.streams)
        vb = mpeg_b200.VideoBatch(ctx, streams, threads=threads, validate=False)
        frames = 0
        ctx.sync()
        t0 = time.perf_counter()
        while True:
            has, buf, _ = vb.step()
            if not has.any():
                break
            frames += int(has.sum())
        ctx.sync()
        dt = time.perf_counter() - t0
        res["gpu"] = {"frames_per_sec": frames / dt, "frames": frames, "seconds": dt, "launches": ctx.launch_count,
                      "note": "lock-step batch: parse (host threads) -> pinned waves -> H2D -> expand_vlen + plan + fused kernels; frames stay in HBM"}
        vb.close()
        ctx.close()
        if args.device_vlc:
            # the same streams with the slices parsed on the GPU: the host scans start codes and copies compressed bytes
            from mpeg_b200.batch import BatchScanStep
            h = L.mpegb200_video_batch_new(len(streams), threads, None, None)
            for i, d in enumerate(streams):
                assert L.mpegb200_video_batch_set_stream(h, i, d, len(d)) == 0
            st = BatchScanStep()
            dt_scan, k, scanned = 0.0, 0, 0
            while True:
                t0 = time.perf_counter()
                assert L.mpegb200_video_batch_next_scan(h, C.byref(st)) == 0
                t1 = time.perf_counter()
                got = int(np.ctypeslib.as_array(st.has_frame, shape=(len(streams),)).sum())
                if got == 0:
                    break
                k += 1
                if k > 1:
                    dt_scan += t1 - t0
                    scanned += sum(st.waves[w].n_pictures for w in range(st.n_waves))
            L.mpegb200_video_batch_free(h)
            ctx = mpeg_b200.Context(0, args.streams)
            ctx.set_kernel_timing(True)
            vb = mpeg_b200.VideoBatch(ctx, streams, threads=threads, validate=False, device_vlc=True)
            frames, parse_ms, steps = 0, [], 0
            ms = C.c_float()
            ctx.sync()
            t0 = time.perf_counter()
            while True:
                has, buf, _ = vb.step()
                if not has.any():
                    break
                frames += int(has.sum())
                steps += 1
                if L.mpegb200_video_bitstream_parse_ms(ctx.h, C.byref(ms)) == 0:
                    parse_ms.append(ms.value)
            ctx.sync()
            dt = time.perf_counter() - t0
            plan, fused = ctx.kernel_times()
            res["device_vlc"] = {"frames_per_sec": frames / dt, "frames": frames, "seconds": dt, "launches": ctx.launch_count,
                                 "flagged_pictures": vb.flagged, "host_steps": vb.host_steps, "steps": steps,
                                 "seconds_in": {"host_scan": vb.t_scan, "submit": vb.t_submit, "waiting_for_flags": vb.t_wait},
                                 "host_scan_pictures_per_sec": scanned / dt_scan if dt_scan else None,
                                 "parse_kernel_ms_per_wave": parse_ms, "decode_ms_per_wave": [float(a + b) for a, b in zip(plan, fused)],
                                 "parse_kernel_pictures_per_sec": (args.streams / (min(parse_ms[1:] or parse_ms) / 1e3)) if parse_ms else None,
                                 "speedup_vs_host_parser_path": (frames / dt) / res["gpu"]["frames_per_sec"],
                                 "note": "host: start-code scan + copy of compressed bytes; device: vlc_parse_kernel (one thread per slice) + "
                                         "vlc_check_kernel + plan + fused kernels; one flag read-back (synchronisation) per wave"}
            vb.close()
            ctx.close()
            if args.resident:
                ctx = mpeg_b200.Context(0, args.streams)
                ctx.set_kernel_timing(True)
                t0 = time.perf_counter()
                vb = mpeg_b200.VideoBatch(ctx, streams, threads=threads, validate=False, device_vlc=True, resident=True)
                ctx.sync()
                t_setup = time.perf_counter() - t0
                frames, parse_ms, steps = 0, [], 0
                t0 = time.perf_counter()
                while True:
                    has, buf, _ = vb.step()
                    if not has.any():
                        break
                    frames += int(has.sum())
                    steps += 1
                    if L.mpegb200_video_bitstream_parse_ms(ctx.h, C.byref(ms)) == 0:
                        parse_ms.append(ms.value)
                ctx.sync()
                dt = time.perf_counter() - t0
                res["device_vlc_resident"] = {"frames_per_sec": frames / dt, "frames": frames, "seconds": dt, "steps": steps,
                                              "setup_seconds": t_setup, "setup": "batch creation incl. upload of every stream to HBM and its start-code index (device)",
                                              "flagged_pictures": vb.flagged, "host_steps": vb.host_steps,
                                              "seconds_in": {"host_scan": vb.t_scan, "submit": vb.t_submit, "waiting_for_flags": vb.t_wait},
                                              "parse_kernel_ms_per_wave_median": sorted(parse_ms)[len(parse_ms) // 2] if parse_ms else None,
                                              "speedup_vs_host_parser_path": (frames / dt) / res["gpu"]["frames_per_sec"]}
                vb.close()
                ctx.close()
    print(json.dumps(res))


if __name__ == "__main__":
    main()
