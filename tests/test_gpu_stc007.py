"""Parity of the CUDA path (through the C ABI) with the oracle on the same seeded inputs.  Run on the B200 box."""
import numpy as np
import pytest

from oracle import oraclebind as O
from sdvpcmdecoder_b200 import synth, capi
from sdvpcmdecoder_b200.capi import LINE_REC, LINE_AUX, BLOCK_REC
from tests import util
from tests.test_hostemu import _random_lines

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import torch
    from sdvpcmdecoder_b200 import operators
    assert torch.cuda.is_available()
    h = capi.Handle(0)
    return h, operators, torch


def _decode(ctx, luma, mode=2, dup=True):
    h, ops, torch = ctx
    v2d = ops.VideoToDigital(h)
    v2d.setBinarizationMode(mode)
    v2d.setCheckLineDup(dup)
    recs, aux = v2d.doBinarize(torch.from_numpy(np.ascontiguousarray(luma)).cuda(), want_aux=True)
    torch.cuda.synchronize()
    return ops.records_to_numpy(recs, LINE_REC), ops.records_to_numpy(aux, LINE_AUX), v2d.stats(), recs


def _check(ctx, luma, mode=2, dup=True):
    o = O.v2d_stc007(mode, luma, dup)
    rec, aux, st, _ = _decode(ctx, luma, mode, dup)
    bad = util.compare_line_records(o, rec, aux)
    assert not bad, bad
    return o, st


def test_clean_tape_all_fields(ctx):
    o, st = _check(ctx, synth.make_stc007(6)["luma"])
    assert st["frames_skipped"] == 5 and st["lines_chain"] == 576
    assert (o["flags"] & 1).mean() > 0.99


def test_clean_tape_ntsc_and_unaligned_width(ctx):
    _check(ctx, synth.make_stc007(4, pal=False)["luma"])
    t = synth.make_stc007(3, pal=False, width=722, x0=15, x1=709)      # stride not a multiple of 16: no bulk-copy path
    _check(ctx, t["luma"])


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_damaged_tape_all_modes(ctx, mode):
    luma = synth.damage_stc007(synth.make_stc007(2)["luma"], seed=100 + mode)
    _check(ctx, luma, mode=mode)


def test_control_block_and_no_dup(ctx):
    _check(ctx, synth.make_stc007(3, control_block=True)["luma"])
    _check(ctx, synth.damage_stc007(synth.make_stc007(2)["luma"], seed=9), dup=False)
    _check(ctx, synth.make_stc007(3)["luma"], dup=False)


def test_sparse_damage_hands_back_to_bulk(ctx):
    luma = synth.make_stc007(12)["luma"].copy()
    luma[3, 200:203] = synth.damage_stc007(luma[3:4, 200:203].copy(), seed=1, jitter=True)[0]
    luma[7, 50] = 0
    luma[7, 301, 300:500] = 255
    x = luma[9, 400].astype(np.float32)
    luma[9, 400] = np.clip((x - 16) * 0.8 + 30, 0, 255).astype(np.uint8)
    luma[9, 398] = 0
    o, st = _check(ctx, luma)
    assert st["frames_skipped"] >= 6


def test_noise_and_silence(ctx):
    t = synth.make_stc007(3)
    _check(ctx, synth.damage_stc007(t["luma"], seed=3, jitter=False, dropout_frac=0.0, marker_kill_frac=0.0))
    # digital silence: duplicate-line rule must not fire on almost-silent lines; blank frame = no PCM at all
    luma = t["luma"].copy()
    luma[1] = 16
    _check(ctx, luma)
    _check(ctx, np.full((2, 576, 720), 16, np.uint8))
    _check(ctx, np.zeros((1, 576, 720), np.uint8))


def test_empty_and_bad_arguments(ctx):
    h, ops, torch = ctx
    v2d = ops.VideoToDigital(h)
    out = v2d.doBinarize(torch.zeros((0, 576, 720), dtype=torch.uint8, device="cuda"))
    assert out.shape[0] == 0
    with pytest.raises(capi.SdvError) as e:
        v2d.doBinarize(torch.zeros((1, 576, 100), dtype=torch.uint8, device="cuda"))       # shorter than one PCM line
    assert e.value.code == capi.SDV_ERR_ARG
    v2d.setPCMType(7)
    with pytest.raises(capi.SdvError) as e:
        v2d.doBinarize(torch.zeros((1, 480, 720), dtype=torch.uint8, device="cuda"))
    assert e.value.code == capi.SDV_ERR_UNSUPPORTED


@pytest.mark.parametrize("res_mode", [0, 1, 2, 3])
@pytest.mark.parametrize("pq", [(True, True), (True, False), (False, False)])
def test_deinterleave_blocks(ctx, res_mode, pq):
    h, ops, torch = ctx
    lines = _random_lines(5000, seed=res_mode * 7 + pq[0] * 2 + pq[1], p_bad=0.06, burst=True)
    ob = O.deint_stc007(lines["words"][:, :8], (lines["flags"] & 3).astype(np.uint8), res_mode, False, True, pq[0], pq[1])
    d = ops.STC007Deinterleaver(h)
    d.setResMode(res_mode); d.setPCorrection(pq[0]); d.setQCorrection(pq[1])
    blocks, samples, flags = d.processBlocks(torch.from_numpy(lines.view(np.uint8).reshape(-1, 32)).cuda())
    torch.cuda.synchronize()
    bad = util.compare_blocks(ob, ops.records_to_numpy(blocks, BLOCK_REC), samples.cpu().numpy(), flags.cpu().numpy())
    assert not bad, bad


def test_broken_block_windows(ctx):
    """The window kernels against the sequential formulation of the same rule (host build of the device code) on random lines.
    The rule itself is pinned by the reference: tests/test_gpu_stc007_stitch.py compares tapes that open such windows with the
    reference pipeline's PCMSamplePair stream and block taps."""
    h, ops, torch = ctx
    lines = _random_lines(6000, seed=77, p_bad=0.03)
    eb, es, ef = util.emu_deint(lines, 0, False, True, True, True, broken_mask_dur=128)
    assert (eb["flags"] & capi.BF_UNSAFE).any()
    d = ops.STC007Deinterleaver(h)
    d.broken_mask_dur = 128
    blocks, samples, flags = d.processBlocks(torch.from_numpy(lines.view(np.uint8).reshape(-1, 32)).cuda())
    torch.cuda.synchronize()
    gb = ops.records_to_numpy(blocks, BLOCK_REC)
    assert np.array_equal(gb, eb)
    assert np.array_equal(samples.cpu().numpy(), es) and np.array_equal(flags.cpu().numpy(), ef)


def _expected_samples(tape, n_blocks, lead_in=80):
    src = np.arange(n_blocks) - (lead_in - tape["j0"])
    return src


def test_tape_to_samples_matches_source_audio(ctx):
    """Encode -> decode round trip at a size the CPU cannot check line by line: every block flagged valid must equal
    the source audio, and all blocks away from the tape ends must be valid (P fixes the 6 uncaptured lines per field)."""
    h, ops, torch = ctx
    tape = synth.make_stc007(40)
    luma = torch.from_numpy(tape["luma"]).cuda()
    v2d = ops.VideoToDigital(h)
    recs = v2d.doBinarize(luma)
    st = ops.STC007DataStitcher(h)
    _, samples, flags = st.doFrameReassemble(recs, 40, 576)
    torch.cuda.synchronize()
    s, f = samples.cpu().numpy(), flags.cpu().numpy()
    src = _expected_samples(tape, len(s))
    ok = (f & capi.SF_BLOCK_OK).all(axis=1)
    inside = (src >= 0) & (src < tape["audio"].shape[0])
    exp = (tape["audio"][src[inside]] << 2).astype(np.uint16).view(np.int16)
    assert np.array_equal(s[inside][ok[inside]], exp[ok[inside]])
    core = inside & (np.arange(len(s)) > 300) & (np.arange(len(s)) < len(s) - 300)
    assert ok[core].all()
    # host-buffer entry point gives the same stream
    s2, f2, r2 = ops.decode_tape_host(h, tape["luma"], want_recs=True)
    assert np.array_equal(s2, s) and np.array_equal(f2, f)
    assert np.array_equal(r2, ops.records_to_numpy(recs, LINE_REC))


def test_reference_pipeline_golden(ctx):
    """Sample stream of the UNMODIFIED reference pipeline (tests/golden/stc007_pipeline_pal.npz, generated by
    tests/golden/make_golden.py from oracle/_ref) against the product's stream."""
    import os
    h, ops, torch = ctx
    path = os.path.join(os.path.dirname(__file__), "golden", "stc007_pipeline_pal.npz")
    g = np.load(path)
    tape = synth.make_stc007(int(g["n_frames"]), seed=int(g["seed"]))
    s, f, _ = ops.decode_tape_host(h, tape["luma"])
    lo, hi = int(g["first_pair"]), int(g["first_pair"]) + len(g["l"])
    got = s.reshape(-1, 2)[lo:hi]
    gf = f.reshape(-1, 2)[lo:hi]
    assert np.array_equal(got[:, 0], g["l"]) and np.array_equal(got[:, 1], g["r"])
    assert np.array_equal(gf[:, 0], g["flags_l"] & 7) and np.array_equal(gf[:, 1], g["flags_r"] & 7)


def test_deinterleave_ignore_crc(ctx):
    from tests.test_hostemu import _lines_with_data_flags
    h, ops, torch = ctx
    lines, crc_ok = _lines_with_data_flags(3000, 23)
    ob = O.deint_stc007(lines["words"][:, :8], crc_ok, 0, True, False, True, True)
    d = ops.STC007Deinterleaver(h)
    d.setIgnoreCRC(True); d.setForcedErrorCheck(False)
    blocks, samples, flags = d.processBlocks(torch.from_numpy(lines.view(np.uint8).reshape(-1, 32)).cuda())
    torch.cuda.synchronize()
    bad = util.compare_blocks(ob, ops.records_to_numpy(blocks, BLOCK_REC), samples.cpu().numpy(), flags.cpu().numpy())
    assert not bad, bad


def test_segment_mode_equals_reference_per_segment(ctx):
    """chain_segments = S decodes the tape as S independent files in one launch: each segment must equal the oracle run
    on that piece of tape (all fields)."""
    h, ops, torch = ctx
    luma = synth.damage_stc007(synth.make_stc007(9, seed=21)["luma"], seed=33)
    v2d = ops.VideoToDigital(h)
    v2d.chain_segments = 4
    recs, aux = v2d.doBinarize(torch.from_numpy(luma).cuda(), want_aux=True)
    torch.cuda.synchronize()
    rec, ax = ops.records_to_numpy(recs, LINE_REC), ops.records_to_numpy(aux, LINE_AUX)
    for s in range(4):
        a, b = s * 9 // 4, (s + 1) * 9 // 4
        o = O.v2d_stc007(2, luma[a:b], True)
        bad = util.compare_line_records(o, rec[a * 576:b * 576], ax[a * 576:b * 576])
        assert not bad, (s, bad)


@pytest.mark.parametrize("seed", range(6))
def test_random_damage_sweep_all_fields(ctx, seed):
    """Differential fuzz of the chain path: different damage mixes, all record fields against the oracle."""
    rng = np.random.RandomState(1000 + seed)
    base = synth.make_stc007(3, seed=200 + seed, pal=bool(seed % 2), control_block=bool(seed % 3 == 0))["luma"]
    luma = synth.damage_stc007(base, seed=300 + seed, sigma=float(rng.choice([0, 6, 12, 20])), jitter=bool(seed % 2 == 0),
                               blur=bool(seed % 3 != 1), dropout_frac=float(rng.choice([0.0, 0.02, 0.1])),
                               marker_kill_frac=float(rng.choice([0.0, 0.01, 0.05])))
    _check(ctx, luma, mode=int(rng.choice([1, 2, 3])))


def test_warm_start_never_changes_results(ctx):
    """The speculative bulk launch with the previous call's presets is scheduling only: repeated decodes (guess right),
    a tape with other levels/geometry (guess wrong), and a damaged tape after a clean one all equal the oracle."""
    h, ops, torch = ctx
    a = synth.make_stc007(5, seed=61)["luma"]
    b = synth.make_stc007(5, seed=62, x0=20, x1=700, black=30, white=180)["luma"]
    c = synth.damage_stc007(a, seed=63)
    for luma in (a, a, b, b, a, c, a):
        _check(ctx, luma)


@pytest.mark.parametrize("width", [352, 1024, 1440, 1920])
def test_frame_widths(ctx, width):
    # pixel-per-bit dependent constants of the marker search / AGC windows, and the bulk ring sized by the row pitch
    m = width / 720.0
    t = synth.make_stc007(3, seed=width, width=width, x0=int(14 * m), x1=width - int(14 * m))
    _check(ctx, t["luma"])
    luma = synth.damage_stc007(t["luma"][:2], seed=width + 1, sigma=6.0, dropout_frac=0.03)
    _check(ctx, luma)



def test_first_frame_hook(ctx):
    # sdv_bin_on_first_frame: called exactly once per decode, with frame 0's records final, results unchanged
    h, ops, torch = ctx
    luma = torch.from_numpy(synth.make_stc007(5, seed=21)["luma"]).cuda()
    v2d = ops.VideoToDigital(h)
    ref = v2d.doBinarize(luma).clone()
    H = luma.shape[1]
    for _ in range(3):          # the later calls take the warm path (hook fires beside the bulk pass)
        seen = []
        out = torch.zeros_like(ref)
        got = v2d.doBinarize(luma, out=out, on_first_frame=lambda: seen.append(out[:H].clone()))
        torch.cuda.synchronize()
        assert len(seen) == 1
        assert torch.equal(seen[0], ref[:H]) and torch.equal(got, ref)
    # no hook left behind
    seen = []
    v2d.doBinarize(luma)
    assert not seen


def test_relay_mode_is_the_single_chain(ctx):
    """A tape whose chain never settles (BASELINE config 4) is decoded by many chains at once after the first 64 frames, each
    kept only if it provably started from the true chain state: all record fields equal the oracle's single sequential chain,
    and the run with relay mode switched off."""
    h, ops, torch = ctx
    luma = synth.damage_stc007(synth.make_stc007(112, seed=71)["luma"], seed=4567)
    o = O.v2d_stc007(2, luma, True)
    rec, aux, st, _ = _decode(ctx, luma)
    assert (st["reserved"] >> 16) >= 8, st          # relay mode ran: pieces in the high half, pieces decoded again in the low half
    bad = util.compare_line_records(o, rec, aux)
    assert not bad, (bad, st)
    v2d = ops.VideoToDigital(h)
    v2d.relay = False
    r2 = v2d.doBinarize(torch.from_numpy(luma).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(ops.records_to_numpy(r2, LINE_REC), rec)
    # a piece whose guessed start state is wrong is decoded again from the true one: a jump of the data coordinates in mid-tape
    # (another capture spliced in) makes the frame medians of the first pass differ from what the chain remembers
    t2 = synth.make_stc007(112, seed=72, x0=22, x1=698)["luma"]
    mix = luma.copy()
    mix[80:] = synth.damage_stc007(t2, seed=4568)[80:]
    o = O.v2d_stc007(2, mix, True)
    rec, aux, st, _ = _decode(ctx, mix)
    bad = util.compare_line_records(o, rec, aux)
    assert not bad, (bad, st)


def test_batches_continue_the_chain(ctx):
    """sdv_bin_config.reserved[3]: a tape fed in batches keeps its chain state (presets, coordinate histories) from call to
    call -- the records equal the single call's and the oracle's, on a clean tape (bulk pass), a sparsely damaged one and a
    config-4 one (relay mode in the later batches)."""
    h, ops, torch = ctx
    clean = synth.make_stc007(14, seed=81)["luma"]
    sparse = clean.copy()
    sparse[4, 100:108] = synth.damage_stc007(sparse[4:5, 100:108].copy(), seed=5)[0]
    sparse[9, 301, 200:400] = 255
    heavy = synth.damage_stc007(synth.make_stc007(60, seed=82)["luma"], seed=4567)
    for luma, cuts in ((clean, [0, 1, 6, 14]), (sparse, [0, 5, 9, 10, 14]), (heavy, [0, 7, 20, 60])):
        o = O.v2d_stc007(2, luma, True)
        v2d = ops.VideoToDigital(h)
        parts = []
        for i in range(len(cuts) - 1):
            r, a = v2d.doBinarize(torch.from_numpy(np.ascontiguousarray(luma[cuts[i]:cuts[i + 1]])).cuda(), want_aux=True, continue_file=(i > 0))
            torch.cuda.synchronize()
            parts.append((ops.records_to_numpy(r, LINE_REC), ops.records_to_numpy(a, LINE_AUX)))
        rec = np.concatenate([p[0] for p in parts])
        aux = np.concatenate([p[1] for p in parts])
        bad = util.compare_line_records(o, rec, aux)
        assert not bad, (cuts, bad)
    with pytest.raises(capi.SdvError):
        ops.VideoToDigital(capi.Handle(0)).doBinarize(torch.from_numpy(clean[:2]).cuda(), continue_file=True)


def test_lazy_verification(ctx):
    """sdv_bin_config.reserved[2] bit 1 + sdv_bin_decode_verify: the decode call returns before the bulk pass has confirmed that it
    took every frame.  Clean tape: the records stand (verify() False).  A tape whose first frame confirms the warm presets but
    whose later frames are damaged: verify() decodes it again (True) and the records equal the oracle's.  A decode call while a
    verification is pending is refused."""
    h, ops, torch = ctx
    a = synth.make_stc007(6, seed=71)["luma"]
    d = a.copy()
    d[3:] = synth.damage_stc007(a[3:], seed=72)
    v2d = ops.VideoToDigital(h)
    ta, td = torch.from_numpy(a).cuda(), torch.from_numpy(d).cuda()
    ref_a = O.v2d_stc007(capi.MODE_NORMAL, a, True)
    ref_d = O.v2d_stc007(capi.MODE_NORMAL, d, True)
    v2d.doBinarize(ta)                                  # warms the handle
    recs = v2d.doBinarize(ta, lazy=True)
    with pytest.raises(capi.SdvError):
        v2d.doBinarize(ta)
    assert v2d.verify() is False
    torch.cuda.synchronize()
    got = ops.records_to_numpy(recs, LINE_REC)
    assert np.array_equal(got["words"], ref_a["words"]) and np.array_equal(got["flags"] & 7, ref_a["flags"] & 7)
    recs = v2d.doBinarize(td, lazy=True)                # frame 0 equals the clean tape's: the warm start hits, frames 3.. are not clean
    assert v2d.verify() is True
    torch.cuda.synchronize()
    got = ops.records_to_numpy(recs, LINE_REC)
    assert np.array_equal(got["words"], ref_d["words"]) and np.array_equal(got["flags"] & 7, ref_d["flags"] & 7)
    assert v2d.verify() is False                        # nothing pending
    recs = v2d.doBinarize(ta, lazy=True)                # cold handle after the redo: the call is not lazy in effect, verify is a no-op
    assert v2d.verify() is False


def test_fused_deinterleave_equals_separate_pass(ctx):
    """sdv_stc007_fuse_next_decode: the bulk pass finishes the blocks that lie inside a frame; samples and flags equal the separate
    deinterleave pass -- cold and warm handle, lazy verification, a tape with BROKEN blocks (lines with a valid CRC but foreign
    words: countdown windows across fused and left-over blocks), NTSC, a shard with a halo and no lead-in; a damaged tape falls
    back to the separate pass."""
    h, ops, torch = ctx

    def run(luma, fuse, pal=True, lead_in=80, halo=None, lazy=False):
        v2d = ops.VideoToDigital(h)
        st = ops.STC007DataStitcher(h)
        st.setVideoStandard(ops.VID_PAL if pal else ops.VID_NTSC)
        st.lead_in = lead_in
        n, H = luma.shape[0], luma.shape[1]
        nb = st.block_count(n)
        smp = torch.full((nb, 6), 12345, dtype=torch.int16, device="cuda")
        fl = torch.full((nb, 6), 99, dtype=torch.uint8, device="cuda")
        t = torch.from_numpy(luma).cuda()
        out = []
        for rep in range(2):            # cold, then warm
            smp.fill_(12345); fl.fill_(99)
            if fuse:
                st.fuseWithNextDecode(smp, fl)
            recs = v2d.doBinarize(t, lazy=lazy)
            st.doFrameReassemble(recs, n, H, samples=smp, flags=fl, halo=halo)
            if lazy:
                assert v2d.verify() is False
            torch.cuda.synchronize()
            out.append((smp.cpu().numpy().copy(), fl.cpu().numpy().copy(), v2d.stats()))
        return out

    base = synth.make_stc007(9, seed=81)["luma"]
    broken = base.copy()
    broken[3, 200] = base[5, 200]           # valid CRC, foreign words: every block through this line fails its parity check
    broken[6, 570] = base[2, 570]           # ... near the end of a frame (blocks left to the separate pass)
    broken[7, 101] = base[1, 101]
    ntsc = synth.make_stc007(7, seed=82, pal=False)["luma"]
    for name, luma, kw in (("clean", base, {}), ("broken", broken, {}), ("broken_lazy", broken, {"lazy": True}), ("ntsc", ntsc, {"pal": False}),
                           ("shard", base[2:7], {"lead_in": 0, "halo": True})):
        kw = dict(kw)
        if kw.pop("halo", False):
            r = ops.VideoToDigital(capi.Handle(0)).doBinarize(torch.from_numpy(base[7:9]).cuda())
            kw["halo"] = r[:112].clone()
        a, b = run(luma, False, **kw), run(luma, True, **kw)
        for (s0, f0, _), (s1, f1, st1) in zip(a, b):
            assert np.array_equal(s0, s1) and np.array_equal(f0, f1), name
        if name == "broken":
            assert (a[0][1] & 1 == 0).any() and not (a[0][1] == 99).any()
    damaged = synth.damage_stc007(base, seed=83)
    a, b = run(damaged, False), run(damaged, True)
    for (s0, f0, _), (s1, f1, _) in zip(a, b):
        assert np.array_equal(s0, s1) and np.array_equal(f0, f1)
