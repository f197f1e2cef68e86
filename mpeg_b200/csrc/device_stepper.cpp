// device_stepper.cpp -- the lock-step Decode() of many streams with their slices parsed on the device, as ONE host call
// (include/mpegb200_host.h, mpegb200_device_stepper_*).  Pure host code over the two C-ABIs: the scan-mode batch of
// host_parser.cpp (headers, start codes, slice tables) and the device entry points of ctx.cu (mpegb200_video_decode_bitstream,
// flags, host-parsed launches).  This is the control flow a cgo binding would otherwise write in Go:
//   1. streams whose step the host parsed itself (stale coefficients pending) run their launches;
//   2. wave by wave: tables (and bytes, unless the streams are resident) to the device, parse + decode kernels;
//   3. while the device works on the step's last wave the host scans the NEXT step (scan-ahead);
//   4. the wave's flags come back; a flagged picture's step is finished by the host parser with the reference's serial
//      semantics (the scan made ahead is withdrawn first), and the stream's pictures in later waves of the step are void.
#include <chrono>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/mpegb200_host.h"

struct mpegb200_device_stepper {
    mpegb200_ctx* ctx = nullptr;
    mpegb200_video_batch* batch = nullptr;
    int first = 0, n = 0;
    bool scan_ahead = true;
    struct Scan {                      // one scan step: the batch's structure plus copies of what the next scan overwrites
        bool valid = false;
        mpegb200_batch_scan_step st;
        std::vector<int> has, buf;
        std::vector<double> time;
    } cur, ahead;
    std::vector<int> flags;
    std::vector<uint8_t> done;
    mpegb200_device_stepper_stats stats;
};

namespace {

using Clock = std::chrono::steady_clock;
double seconds(Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double>(b - a).count(); }

int scan(mpegb200_device_stepper* s, mpegb200_device_stepper::Scan& out) {
    const auto t0 = Clock::now();
    const int rc = mpegb200_video_batch_next_scan(s->batch, &out.st);
    s->stats.seconds_host_scan += seconds(t0, Clock::now());
    if (rc != 0) return rc;
    s->n = out.st.n_streams;
    out.has.assign(out.st.has_frame, out.st.has_frame + s->n);
    out.buf.assign(out.st.frame_buf, out.st.frame_buf + s->n);
    out.time.assign(out.st.time, out.st.time + s->n);
    out.valid = true;
    return 0;
}

// the launches of a host-parsed step for context stream `id`
int run_launches(mpegb200_device_stepper* s, const mpegb200_video_step& step, int id) {
    for (int i = 0; i < step.n_launches; i++) {
        const mpegb200_launch& L = step.launches[i];
        if (L.n_mb == 0) continue;
        mpegb200_picture pic = L.picture;
        pic.stream = id;
        int rc;
        if (step.vlen_launches) {
            const mpegb200_launch_vlen& V = step.vlen_launches[i];
            rc = mpegb200_video_decode_pictures_vlen(s->ctx, 1, &pic, L.n_mb, step.mbs + L.first_mb, L.n_blocks, step.vlen_headers + L.first_block,
                                                     step.vlen_chunk_offsets + V.first_chunk, step.vlen_payload + V.payload_offset, V.payload_bytes);
        } else {
            rc = mpegb200_video_decode_pictures(s->ctx, 1, &pic, L.n_mb, step.mbs + L.first_mb, L.n_blocks, step.coeffs + (size_t)L.first_block * 64);
        }
        if (rc != 0) return rc;
    }
    return mpegb200_sync_uploads(s->ctx);   // the parser re-uses these arrays on its next call
}

int step_impl(mpegb200_device_stepper* s, int* has_frame, int* frame_buf, double* time) {
    if (s->ahead.valid) {
        std::swap(s->cur, s->ahead);
        s->ahead.valid = false;
    } else if (int rc = scan(s, s->cur)) {
        return rc;
    }
    mpegb200_device_stepper::Scan& cur = s->cur;
    const int n = s->n;
    memcpy(has_frame, cur.has.data(), sizeof(int) * (size_t)n);
    memcpy(frame_buf, cur.buf.data(), sizeof(int) * (size_t)n);
    memcpy(time, cur.time.data(), sizeof(double) * (size_t)n);
    for (int j = 0; j < cur.st.n_host; j++) {
        if (int rc = run_launches(s, cur.st.host_steps[j], cur.st.host_index[j] + s->first)) return rc;
        s->stats.host_steps++;
    }
    s->done.assign((size_t)n, 0);
    for (int w = 0; w < cur.st.n_waves; w++) {
        const mpegb200_vlc_wave& W = cur.st.waves[w];
        if (W.n_pictures == 0) continue;
        // the wave's tables are the batch's own (pinned) arrays of this step: void pictures and stream ids are patched in place
        mpegb200_vlc_picture* pics = const_cast<mpegb200_vlc_picture*>(W.pics);
        for (int k = 0; k < W.n_pictures; k++) {
            if (s->done[(size_t)pics[k].stream]) pics[k].type = 0;
            pics[k].stream += s->first;
        }
        auto t0 = Clock::now();
        if (int rc = mpegb200_video_decode_bitstream(s->ctx, W.n_pictures, pics, W.n_slices, W.slices, W.bitstream, W.bitstream_bytes, W.quant,
                                                     W.n_quant, W.n_mb_slots))
            return rc;
        s->stats.seconds_submit += seconds(t0, Clock::now());
        s->stats.waves++;
        // a guess: no picture of this wave will flag.  If one does, the scan is withdrawn below.
        if (s->scan_ahead && w == cur.st.n_waves - 1)
            if (int rc = scan(s, s->ahead)) return rc;
        t0 = Clock::now();
        s->flags.resize((size_t)W.n_pictures);
        const int bad = mpegb200_video_bitstream_flags(s->ctx, s->flags.data(), W.n_pictures);
        s->stats.seconds_waiting += seconds(t0, Clock::now());
        if (bad < 0) return bad;
        for (int k = 0; bad > 0 && k < W.n_pictures; k++) {
            if (!s->flags[(size_t)k]) continue;
            const int index = pics[k].stream - s->first;
            if (s->done[(size_t)index]) continue;
            if (s->ahead.valid) {   // the parsers go back to where they stood after this step's scan
                if (int rc = mpegb200_video_batch_unscan(s->batch)) return rc;
                s->ahead.valid = false;
                s->stats.withdrawn_scans++;
            }
            mpegb200_video_step tail;
            if (int rc = mpegb200_video_batch_redo(s->batch, index, W.step_picture[k], &tail)) return rc;
            if (int rc = run_launches(s, tail, index + s->first)) return rc;
            has_frame[index] = tail.has_frame;
            frame_buf[index] = tail.frame_buf;
            time[index] = tail.time;
            s->done[(size_t)index] = 1;
            s->stats.flagged_pictures++;
        }
    }
    s->stats.steps++;
    return 0;
}

}  // namespace

extern "C" {

mpegb200_device_stepper* mpegb200_device_stepper_new(mpegb200_ctx* ctx, mpegb200_video_batch* batch, int first_stream, int scan_ahead) {
    if (!ctx || !batch || first_stream < 0) return nullptr;
    auto* s = new (std::nothrow) mpegb200_device_stepper();
    if (!s) return nullptr;
    s->ctx = ctx;
    s->batch = batch;
    s->first = first_stream;
    s->scan_ahead = scan_ahead != 0;
    memset(&s->stats, 0, sizeof(s->stats));
    return s;
}

void mpegb200_device_stepper_free(mpegb200_device_stepper* s) { delete s; }

int mpegb200_device_stepper_step(mpegb200_device_stepper* s, int* has_frame, int* frame_buf, double* time) {
    if (!s || !has_frame || !frame_buf || !time) return MPEGB200_EINVAL;
    try {
        return step_impl(s, has_frame, frame_buf, time);
    } catch (...) {
        return MPEGB200_ENOMEM;
    }
}

int mpegb200_device_stepper_drop_scan_ahead(mpegb200_device_stepper* s) {
    if (!s) return MPEGB200_EINVAL;
    if (s->ahead.valid) {
        s->ahead.valid = false;
        return mpegb200_video_batch_unscan(s->batch);
    }
    return 0;
}

int mpegb200_device_stepper_get_stats(mpegb200_device_stepper* s, mpegb200_device_stepper_stats* out) {
    if (!s || !out) return MPEGB200_EINVAL;
    *out = s->stats;
    return 0;
}

}  // extern "C"
