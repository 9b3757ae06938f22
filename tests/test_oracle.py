"""CPU tests of the oracle: it must reproduce, byte for byte, what the
reference's own scripts printed and wrote (committed golden fixtures, made by
tests/golden/make_golden.py from /root/reference), and - where the reference
tree is present - agree with a fresh execution of those scripts."""

import os

import numpy as np
import pytest

import cases
from conftest import load_golden, run_oracle
from oracle import ref_exec, distances as OD

ALL = sorted(cases.CASES)


def _argv(name, tmp, out):
    rpath, feadir, sha, _ = cases.materialise(name, str(tmp))
    return [rpath, feadir, '-o', out] + cases.CASES[name][3], sha


@pytest.mark.parametrize('name', ALL)
def test_oracle_reproduces_reference_golden(name, tmp_path):
    kind, variant, _, _ = cases.CASES[name]
    gold = load_golden(name)
    out = str(tmp_path / 'out.recipe')
    argv, sha = _argv(name, tmp_path, out)
    assert sha == gold['frames_sha256'], 'synthetic generator drifted'
    if 'raises' in gold:
        with pytest.raises(ValueError) as e:
            run_oracle(kind, variant, argv + ['--sw-bic', 'strict'])
        assert 'ValueError: %s' % e.value == gold['raises']
        return
    stdout, _ = run_oracle(kind, variant, argv)
    assert open(out).read() == gold['recipe']
    assert stdout.replace(str(tmp_path), '<TMP>') == gold['stdout']


@pytest.mark.skipif(not ref_exec.available(), reason='reference tree not present')
@pytest.mark.parametrize('name', ['gw_bic_f125', 'gw_bic_multi', 'sw_glr', 'merge_bic', 'cl1_hi_bic',
                                  'cl2_hi_bic', 'cl1_in_bic'])
def test_oracle_matches_live_reference(name, tmp_path):
    kind, variant, _, _ = cases.CASES[name]
    script = {('cd', 0): 'spk-change-detection.py', ('cl', 1): 'spk-clustering.py',
              ('cl', 2): 'spk-clustering2.py'}[(kind, variant)]
    out_o, out_r = str(tmp_path / 'o.recipe'), str(tmp_path / 'r.recipe')
    argv, _ = _argv(name, tmp_path, out_o)
    so, _ = run_oracle(kind, variant, argv)
    sr, _ = ref_exec.run(script, argv[:3] + [out_r] + argv[4:])
    assert open(out_o).read() == open(out_r).read()
    assert so.replace(out_o, 'X') == sr.replace(out_r, 'X')


def test_q5_variants_differ():
    """variant 2 keeps stale distances, so its merge sequence is
    not variant 1's on the same input (a property of the reference worth
    keeping visible)."""
    g1, g2 = load_golden('cl1_hi_q5'), load_golden('cl2_hi_q5')
    m1 = [l for l in g1['stdout'].splitlines() if l.startswith('Merging:')]
    m2 = [l for l in g2['stdout'].splitlines() if l.startswith('Merging:')]
    assert m1 and m2 and m1 != m2
    assert 'Final speakers: 5' in g1['stdout'] and 'Final speakers: 3' in g2['stdout']


def test_kl2_is_the_diagonal_formula():
    """element-wise products make the reference's KL2 a
    diagonal-only formula with float32 sequential means."""
    rng = np.random.default_rng(0)
    a = (rng.standard_normal((300, 39)) * 1.3 + 0.2).astype(np.float32)
    b = (rng.standard_normal((200, 39)) * 0.8 - 0.1).astype(np.float32)
    S1, S2 = np.cov(a, rowvar=0), np.cov(b, rowvar=0)
    P1, P2 = np.linalg.inv(S1), np.linalg.inv(S2)

    def seq_mean(x):
        s = np.zeros(39, dtype=np.float32)
        for row in x:
            s = s + row
        return s / np.float32(x.shape[0])
    delta = (seq_mean(a) - seq_mean(b)).astype(np.float64)
    want = 0.5 * np.sum(np.diag(S1 - S2) * np.diag(P2 - P1)) + 0.5 * np.sum(np.diag(P1 + P2) * delta * delta)
    got = OD.kl2(a, b)
    assert abs(got - want) <= 1e-11 * abs(want)
    # the textbook (matrix-product) KL2 is a different number
    textbook = 0.5 * np.trace((S1 - S2) @ (P2 - P1)) + 0.5 * delta @ (P1 + P2) @ delta
    assert abs(textbook - got) > 1e-3 * abs(got)


def test_bic_memo_quirk():
    """the shared memo makes every later sw/merge BIC reuse the
    first left term."""
    rng = np.random.default_rng(1)
    x = rng.standard_normal((900, 39)).astype(np.float32)
    memo = OD.BicMemo()
    d0 = OD.bic_cd(x[:100], x[100:200], x[:200], 1.3, 0, memo)
    d1 = OD.bic_cd(x[300:500], x[500:700], x[300:700], 1.3, 0, memo)
    own = OD.bic_cd(x[300:500], x[500:700], x[300:700], 1.3, 0, OD.BicMemo())
    assert d0 == OD.bic_cd(x[:100], x[100:200], x[:200], 1.3, 0, OD.BicMemo())
    assert d1 != own and abs((d1 - own) - (0.5 * 200 * np.log(np.linalg.det(np.cov(x[300:500], rowvar=0)))
                                             - memo[0])) < 1e-8
