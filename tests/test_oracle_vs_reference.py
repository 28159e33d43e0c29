"""Live comparison of the oracle (oracle/chain.py, oracle/bs1770.py) with the UNMODIFIED reference imported from
/root/reference through oracle/ref_harness.py (stand-ins only for the three absent I/O packages).  Runs where the reference
tree exists (the build container); skipped on the GPU box, where the committed goldens (tests/golden/*.npz, made from the same
reference by tests/golden/make_golden*.py) carry the pin.  Inputs here are FRESH (not the golden recipes): all eight presets,
both chains, three sample rates -- the oracle must return the reference's samples bit for bit."""
import warnings

import numpy as np
import pytest

from oracle import chain as oc, ref_harness
from mm_b200 import synth

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="reference tree not present (GPU box): goldens carry the pin")

STYLES = ["standard", "edm", "hiphop", "classical", "podcast", "lofi", "house_basic", "dry_vocal"]


@pytest.fixture(scope="module")
def ref():
    return ref_harness.load()


def _v2(ref, x, sr, target, style):
    ch = ref.chain.MasteringChain.default_chain(target_lufs=target, style=style)
    out = ch.process(x, sr, target_lufs=target, style=style)
    return ref.pipeline.apply_output_edge_fade_in(out, sr, fade_ms=6.0)          # routers/mastering.py:583


@pytest.mark.parametrize("k", range(8))
def test_chains_equal_reference_all_presets(ref, k):
    style = STYLES[k]
    sr = (44100, 48000, 96000)[k % 3]
    dur = 3.0 if sr == 96000 else 5.0
    x = synth.numpy_track(40 + k, sr, dur)
    if k == 5:
        x = np.ascontiguousarray(x[:, 0])                                       # one mono case
    target = oc.STYLE_CONFIGS[style]["lufs"]
    r1 = ref.pipeline.run_mastering_pipeline(x.copy(), sr, target_lufs=target, style=style)
    o1 = oc.run_v1(x.copy(), sr, target, style)
    assert r1.dtype == o1.dtype and np.array_equal(r1, o1), (style, "v1", float(np.max(np.abs(r1 - o1))))
    r2 = _v2(ref, x.copy(), sr, target, style)
    o2 = oc.run_v2(x.copy(), sr, target, style)
    assert np.array_equal(r2, o2), (style, "v2", float(np.max(np.abs(r2 - o2))))
    assert ref.pipeline.measure_lufs(r1, sr) == oc.measure_lufs(o1, sr)
    assert ref.true_peak_dbfs(r2, sr) == oc.true_peak_dbfs(o2, sr)


def test_dynamic_eq_default_bands_equal_reference_20s(ref):
    for t, sr in ((6, 44100), (7, 48000), (9, 48000)):
        x = synth.numpy_track(t, sr, 20.0)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert np.array_equal(ref.pipeline.apply_dynamic_eq(x, sr), oc.apply_dynamic_eq(x, sr)), (t, sr)


def test_pro_stages_equal_reference(ref):
    sr = 48000
    x = synth.numpy_track(50, sr, 4.0)
    P = ref.pipeline
    pairs = [
        (P.apply_rumble_filter(x, sr, 80.0), oc.zero_phase(*oc.sg.butter(2, 80.0 / (sr / 2), btype="high"), x.astype(np.float64).T).T.astype(np.float32)),
        (P.apply_transient_designer(x, sr, 1.4, 0.8), oc.apply_transient_designer(x, sr, 1.4, 0.8)),
        (P.apply_maximizer_transient_aware(x, sr, 0.5), oc.apply_maximizer_transient_aware(x, sr, 0.5)),
        (P.apply_high_freq_trim(x, sr), oc.apply_high_freq_trim(x, sr)),
        (P.apply_deesser(x, sr), oc.apply_deesser(x, sr)),
        (P.apply_parallel_compression(x, sr, mix=0.3), oc.apply_parallel_compression(x, sr, mix=0.3)),
        (P.apply_reverb(x, sr, "room", 0.8, 0.2), oc.apply_reverb(x, sr, "room", 0.8, 0.2)),
    ]
    for i, (a, b) in enumerate(pairs):
        assert a.shape == b.shape and float(np.max(np.abs(a.astype(np.float64) - b))) <= 1e-7, i


def test_int16_export_equal_reference(ref):
    sr = 44100
    x = synth.numpy_track(51, sr, 2.0)
    rng = np.random.default_rng(3)
    noise = (rng.random(x.shape) + rng.random(x.shape) - 1.0).astype(np.float32)
    P = ref.pipeline
    orig = P._dither_noise_tpdf
    P._dither_noise_tpdf = lambda shape: noise
    try:
        wav = P.export_audio(x, sr, 2, "wav", dither_type="tpdf")
    finally:
        P._dither_noise_tpdf = orig
    pcm = np.frombuffer(wav[44:], dtype="<i2").reshape(-1, 2)
    assert np.array_equal(pcm, oc.quantize_int16(x, noise))


def test_prebuilt_reference_copy_runs_sourceless_and_equals_the_oracle(tmp_path):
    """oracle/make_ref.py byte-compiles the reference's mastering path into oracle/_ref (what bench.py's CPU legs import on the
    GPU box, where /root/reference does not exist).  Imported SOURCELESS in a fresh process it must return the oracle's samples
    bit for bit, like the tree itself does above."""
    import os
    import subprocess
    import sys
    from oracle import make_ref
    if not os.path.isfile(os.path.join(make_ref.SRC, "app", "pipeline.py")):
        pytest.skip("reference tree not present: nothing to build from")
    out_dir = make_ref.build(verbose=False)
    assert not [f for _, _, fs in os.walk(out_dir) for f in fs if f.endswith(".py")], "no reference source may be copied"
    code = (
        "import sys, numpy as np\n"
        "from oracle import ref_harness, chain as oc\n"
        "from mm_b200 import synth\n"
        "assert ref_harness.available() and not ref_harness.is_source_tree()\n"
        "ref = ref_harness.load()\n"
        "assert ref.pipeline.__file__.endswith('.bin')\n"
        "x = synth.numpy_track(77, 48000, 3.0)\n"
        "r1 = ref.pipeline.run_mastering_pipeline(x.copy(), 48000, target_lufs=-9.0, style='edm')\n"
        "assert np.array_equal(r1, oc.run_v1(x.copy(), 48000, -9.0, 'edm'))\n"
        "ch = ref.chain.MasteringChain.default_chain(target_lufs=-14.0, style='standard')\n"
        "r2 = ref.pipeline.apply_output_edge_fade_in(ch.process(x.copy(), 48000, target_lufs=-14.0, style='standard'), 48000, fade_ms=6.0)\n"
        "assert np.array_equal(r2, oc.run_v2(x.copy(), 48000, -14.0, 'standard'))\n"
        "assert ref.true_peak_dbfs(r2, 48000) == oc.true_peak_dbfs(r2, 48000)\n"
        "print('prebuilt ok')\n")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = {**os.environ, "MM_REFERENCE_BACKEND": out_dir, "PYTHONPATH": repo + os.pathsep + os.path.join(repo, "audio-mastering-web_b200"),
           "NUMBA_CACHE_DIR": str(tmp_path)}
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "prebuilt ok" in r.stdout, r.stderr[-2000:]


@pytest.mark.parametrize("n", [3, 9, 10, 15, 16])
def test_tiny_inputs_equal_reference(ref, n):
    """Inputs of <= padlen samples: scipy's filtfilt raises and the reference's _safe_filtfilt returns lfilter (pipeline.py:36-52).
    The oracle must follow the reference there bit for bit -- it is what the GPU's degrade path is tested against
    (tests/test_gpu_stages.py::test_inputs_not_longer_than_padlen_degrade_to_lfilter)."""
    rng = np.random.default_rng(100 + n)
    sr = 44100
    P = ref.pipeline
    for ch in (2, 1):
        x = (0.3 * rng.standard_normal((n, ch))).astype(np.float32)
        x = x if ch == 2 else x[:, 0]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            pairs = [(P.apply_target_curve(x.copy(), sr), oc.apply_target_curve(x.copy(), sr)),
                     (P.apply_dynamics(x.copy(), sr), oc.apply_dynamics(x.copy(), sr)),
                     (P.apply_final_spectral_balance(x.copy(), sr), oc.apply_final_spectral_balance(x.copy(), sr)),
                     (P.apply_style_eq(x.copy(), sr, "edm"), oc.apply_style_eq(x.copy(), sr, "edm")),
                     (P.apply_harmonic_exciter(x.copy(), sr, 1.2), oc.apply_harmonic_exciter(x.copy(), sr, 1.2))]
        for i, (a, b) in enumerate(pairs):
            assert a.shape == b.shape and np.array_equal(a, b), (n, ch, i)
