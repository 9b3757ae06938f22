"""Feature files of fewer than 39 dimensions (the reference's loader takes the dimension from the file header,
spk-change-detection.py:37-41): frames are zero-padded to the 39 columns the kernels are compiled for and the
padding is an identity block of the matrices that are factorised, so every distance is that of the real
dimensions - against the oracle, which works in the file's own dimension."""

import io

import numpy as np
import pytest

import spkdiar                              # noqa: F401
from conftest import logs_match, run_oracle, run_product
from oracle import distances as OD
from spkdiar import _abi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ctx():
    c = _abi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize('dim', [13, 26, 38])
def test_window_distances_in_the_files_dimension(ctx, dim):
    rec = synth.make_recording(500 + dim, 4000, 3, dim=dim)
    x = rec.frames
    rng = np.random.default_rng(dim)
    a = rng.integers(0, 1500, 12)
    m = a + rng.integers(120, 900, 12)
    b = m + rng.integers(120, 900, 12)
    with ctx.upload(x) as feat:
        assert feat.dim == dim
        for met, ref in ((_abi.BIC, lambda p, q: OD.bic_cd(p, q, np.concatenate((p, q)), 1.3)),
                         (_abi.GLR, OD.glr), (_abi.KL2, OD.kl2)):
            d = feat.score_windows(a, m, b, met, 1.3)
            for k in range(len(a)):
                want = ref(x[a[k]:m[k]], x[m[k]:b[k]])
                assert abs(d[k] - want) <= 1e-9 * max(abs(want), 1000.0), (dim, met, k, d[k], want)


@pytest.mark.parametrize('dim', [13, 26])
def test_both_stages_on_a_13_and_a_26_dimensional_file(tmp_path, ctx, dim):
    rec = synth.make_recording(600 + dim, 9000, 3, dim=dim, turn_lo=3, turn_hi=9)
    rpath, feadir = synth.write_case(str(tmp_path), 't', rec, synth.one_line_recipe('/syn/t.wav', rec))
    og, pg = str(tmp_path / 'o.recipe'), str(tmp_path / 'p.recipe')
    flags = ['-f', '100', '-m', 'gw', '-d', 'BIC', '-w', '1.0', '-st', '3.0', '-dws', '0.1', '-l', '1.0']
    so, _ = run_oracle('cd', 0, [rpath, feadir, '-o', og] + flags)
    sp, _ = run_product('cd', 0, [rpath, feadir, '-o', pg] + flags, ctx)
    assert open(pg).read() == open(og).read()
    assert len(open(pg).read().splitlines()) > 3
    assert logs_match(sp.replace(pg, 'X'), so.replace(og, 'X')) is None
    for variant in (1, 2):
        oc, pc = str(tmp_path / 'oc.recipe'), str(tmp_path / 'pc.recipe')
        so, _ = run_oracle('cl', variant, [og, feadir + '/', '-o', oc, '-f', '100'])
        sp, _ = run_product('cl', variant, [pg, feadir + '/', '-o', pc, '-f', '100'], ctx)
        assert open(pc).read() == open(oc).read()
        assert logs_match(sp.replace(pc, 'X').replace(pg, 'Y'), so.replace(oc, 'X').replace(og, 'Y')) is None


def test_one_dimension_per_device_at_a_time(ctx):
    a = synth.make_recording(1, 600, 2).frames
    b = synth.make_recording(2, 600, 2, dim=13).frames
    with ctx.upload(a) as fa:
        with pytest.raises(_abi.SpkdiarError):
            ctx.upload(b)
        d39 = fa.score_windows([0], [300], [600], _abi.BIC)
    with ctx.upload(b) as fb:                       # no 39-dimensional handle alive any more
        d13 = fb.score_windows([0], [300], [600], _abi.BIC)
    with ctx.upload(a) as fa:                       # and back
        assert fa.score_windows([0], [300], [600], _abi.BIC).tobytes() == d39.tobytes()
    assert np.isfinite(d13).all()
    with pytest.raises(_abi.SpkdiarError):
        ctx.upload(np.zeros((100, 40), dtype=np.float32))
