"""The test-side MPEG-1 writer (tests/mpeg1_writer.py) against both parsers: a stream written from a quantised-domain
description must decode -- through the product's host parser and through the oracle's restatement of the reference
parser -- to exactly the records `expected_records` computes independently (missing 8).  CPU only."""
import numpy as np
import pytest

import mpeg1_writer as mw
import oracle_lib as ol
from test_host_parser import oracle_steps, parser_steps


def write_stream(width, height, pictures, seed, mode, mv_range=32, f_code=2, scale=8):
    rng = np.random.default_rng(seed)
    w = mw.StreamWriter(width, height, quantizer_scale=scale, f_code=f_code)
    lw, lh = w.mb_w * 16, w.mb_h * 16
    specs = []
    for t in pictures:
        mbs = mw.random_picture(rng, w.mb_w, w.mb_h, t, mode, mv_range, lw, lh)
        w.picture(t, mbs)
        specs.append((t, mbs))
    return w, specs


def check_against(steps, w, specs):
    """steps: what a parser made of the stream (display-order steps, each with the launches decoded on the way)."""
    launches = [l for _, _, ls in steps for l in ls]
    assert len(launches) == len(specs)
    for (hdr, mbs, coeffs), (ptype, spec) in zip(launches, specs):
        recs, blocks = mw.expected_records(spec, w.mb_w, ptype, w.q)
        assert hdr[0] == ptype and hdr[4] == len(recs)
        got = [(int(m["mb_row"]), int(m["mb_col"]), int(m["mv_h"]), int(m["mv_v"]), int(m["flags"]), int(m["cbp"])) for m in mbs]
        assert got == recs
        assert np.array_equal(coeffs, blocks)


@pytest.mark.parametrize("size,pictures,mode", [
    ((64, 48), [mw.PIC_I, mw.PIC_P, mw.PIC_P], "natural"),
    ((64, 48), [mw.PIC_I, mw.PIC_P, mw.PIC_B, mw.PIC_B, mw.PIC_P], "natural"),
    ((96, 64), [mw.PIC_I, mw.PIC_P], "dense"),
    ((352, 288), [mw.PIC_I, mw.PIC_P, mw.PIC_B], "natural"),
])
def test_written_streams_parse_to_the_expected_records(size, pictures, mode):
    w, specs = write_stream(size[0], size[1], pictures, seed=size[0] + len(pictures), mode=mode)
    data = w.tobytes()
    check_against(list(parser_steps(data)), w, specs)      # the product's host parser
    check_against(list(oracle_steps(data)), w, specs)      # the oracle's restatement of the reference parser


def test_wide_vectors_and_other_quantiser_scales():
    for f_code, mv_range, scale in ((4, 128, 3), (1, 16, 31), (3, 64, 1)):
        w, specs = write_stream(160, 128, [mw.PIC_I, mw.PIC_P, mw.PIC_B], seed=f_code, mode="natural", mv_range=mv_range, f_code=f_code, scale=scale)
        check_against(list(parser_steps(w.tobytes())), w, specs)


def test_written_stream_decodes_through_the_oracle_decoder():
    """End to end on the CPU: the oracle's full decoder (parse + reconstruct) accepts the stream and returns one frame per
    picture (plus the flush), with the frame size of the sequence header."""
    w, specs = write_stream(64, 48, [mw.PIC_I, mw.PIC_P, mw.PIC_P, mw.PIC_P], seed=9, mode="natural")
    v = ol.VideoOracle(w.tobytes())
    assert v.has_header() and (v.width, v.height) == (64, 48)
    frames = 0
    while v.decode() is not None:
        frames += 1
    assert frames == len(specs)
