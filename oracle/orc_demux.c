/*
 * orc_demux.c -- minimal MPEG-PS splitter (TEST INFRASTRUCTURE, see mpeg_oracle.h).
 * Restates the packet walk of demux.go (HasHeaders :85-138, Decode :473-511,
 * decodePacket :531-568) far enough to separate testdata/test.mpg into its video (0xE0)
 * and first audio (0xC0) elementary streams, so that BASELINE config 1 can be run through
 * the video/audio oracles.  Seeking, durations and PTS bookkeeping are not part of the hot
 * path and are not restated.
 */
#include "mpeg_oracle.h"
#include "orc_bits.h"

#include <stdlib.h>
#include <string.h>

enum { START_PACK = 0xBA, START_SYSTEM = 0xBB, PACKET_PRIVATE = 0xBD, PACKET_AUDIO_1 = 0xC0, PACKET_AUDIO_4 = 0xC3,
       PACKET_VIDEO_1 = 0xE0 };

typedef struct grow {
    uint8_t* p;
    size_t n, cap;
} grow;

static void grow_add(grow* g, const uint8_t* src, size_t n) {
    if (g->n + n > g->cap) {
        g->cap = (g->n + n) * 2 + 4096;
        g->p = (uint8_t*)realloc(g->p, g->cap);
    }
    memcpy(g->p + g->n, src, n);
    g->n += n;
}

/* decodeTime, demux.go:520-529 (value unused here; the bits must still be consumed) */
static void skip_time(orc_bits* b) {
    bits_read(b, 3);
    bits_skip(b, 1);
    bits_read(b, 15);
    bits_skip(b, 1);
    bits_read(b, 15);
    bits_skip(b, 1);
}

int orc_demux_split(const uint8_t* data, size_t len, uint8_t** video, size_t* video_len, uint8_t** audio,
                    size_t* audio_len, int* n_video_packets, int* n_audio_packets) {
    orc_bits b = {data, len, 0, 0};
    grow gv = {0, 0, 0}, ga = {0, 0, 0};
    int nv = 0, na = 0;

    /* pack header, demux.go:91-113 */
    if (bits_find_start_code(&b, START_PACK) == -1 || !bits_has(&b, 64)) return -1;
    if (bits_read(&b, 4) != 0x02) return -1;
    skip_time(&b);
    bits_skip(&b, 1);
    bits_skip(&b, 22);
    bits_skip(&b, 1);
    /* system header, demux.go:116-133 */
    if (bits_find_start_code(&b, START_SYSTEM) == -1 || !bits_has(&b, 56)) return -1;
    bits_skip(&b, 16);
    bits_skip(&b, 24);
    bits_read(&b, 6);
    bits_skip(&b, 5);
    bits_read(&b, 5);

    for (;;) { /* Decode loop, demux.go:500-510 */
        int code = bits_next_start_code(&b);
        if (code == -1) break;
        if (!(code == PACKET_VIDEO_1 || code == PACKET_PRIVATE || (code >= PACKET_AUDIO_1 && code <= PACKET_AUDIO_4)))
            continue;
        /* decodePacket, demux.go:531-568 */
        if (!bits_has(&b, 16 << 3)) break;
        int64_t length = bits_read(&b, 16);
        length -= bits_skip_bytes(&b, 0xff);
        if (bits_read(&b, 2) == 0x01) {
            bits_skip(&b, 16);
            length -= 2;
        }
        int marker = (int)bits_read(&b, 2);
        if (marker == 0x03) {
            skip_time(&b);
            bits_skip(&b, 40);
            length -= 10;
        } else if (marker == 0x02) {
            skip_time(&b);
            length -= 5;
        } else if (marker == 0x00) {
            bits_skip(&b, 4);
            length -= 1;
        } else {
            continue; /* invalid */
        }
        if (length < 0 || !bits_has(&b, length << 3)) break;
        const uint8_t* payload = data + (b.bit_index >> 3);
        if (code == PACKET_VIDEO_1) {
            grow_add(&gv, payload, (size_t)length);
            nv++;
        } else if (code == PACKET_AUDIO_1) {
            grow_add(&ga, payload, (size_t)length);
            na++;
        }
        bits_skip(&b, (int)(length << 3)); /* demux.go:478-485 */
    }
    *video = gv.p;
    *video_len = gv.n;
    *audio = ga.p;
    *audio_len = ga.n;
    if (n_video_packets) *n_video_packets = nv;
    if (n_audio_packets) *n_audio_packets = na;
    return 0;
}
