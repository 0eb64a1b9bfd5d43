"""Short runs of the randomised campaigns (tests/fuzz/) inside the test suites: the first seconds of seeds whose long runs are on
record (profiles/r02_fuzz_*.txt), so a suite run is reproducible; RACC_FUZZ_SEED=<n> draws other scenes. What the campaigns
found: DESIGN.md section 3."""
import os
import sys

import pytest

import oracle

FUZZ = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fuzz")
sys.path.insert(0, FUZZ)
SEED = os.environ.get("RACC_FUZZ_SEED", "7")


def _run(module, argv):
    mod = __import__(module)
    saved = sys.argv
    sys.argv = [module] + argv
    try:
        return mod.main()
    finally:
        sys.argv = saved


@pytest.mark.gpu
def test_randomised_campaign_on_the_gpu(capsys):
    """Engine vs checker on random scene families x adversarial rays x launch shapes x HOST / DEVICE streams x both builders x the
    renderers, 20 seconds of it (about 60 M rays on a B200)."""
    torch = pytest.importorskip("torch")
    assert torch.cuda.is_available()
    rc = _run("fuzz_gpu", ["--seconds", "20", "--seed", SEED])
    out = capsys.readouterr().out
    assert rc == 0, f"seed {SEED}:\n{out[-3000:]}"


@pytest.mark.skipif(not oracle.have_ref_kernel(), reason="oracle/_ref/libkernel_ref.so not built (needs /root/reference)")
def test_randomised_checker_against_reference_kernel_source(capsys):
    rc = _run("fuzz_oracle_cpu", ["--seconds", "8", "--seed", SEED])
    out = capsys.readouterr().out
    assert rc == 0, f"seed {SEED}:\n{out[-3000:]}"


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref/libracc_ref.so not built (needs /root/reference)")
def test_randomised_builder_against_reference_builder(capsys):
    rc = _run("fuzz_builder_cpu", ["--seconds", "8", "--seed", SEED])
    out = capsys.readouterr().out
    assert rc == 0, f"seed {SEED}:\n{out[-3000:]}"
