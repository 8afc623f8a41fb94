"""Row a1 / b of SURVEY.md §8: the reference's OWN callers, unmodified, on the drop-in classes.

Each case runs tests/run_reference_callers.py in a subprocess (release_server.py flips process-wide torch state at
import).  Needs the reference checkout (/root/reference: the build container; absent on the GPU box, where the
bit-equal harness stand-ins are used instead — tests/test_server_loop_gpu.py, tests/test_pipeline_gpu.py)."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

from tests import ref_env

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.skipif(not ref_env.available(), reason="reference checkout not present")


def run(*args):
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "run_reference_callers.py"), *args],
                       capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("mode", ["keep", "reencode"])
def test_unmodified_generation_session_on_the_dropins(mode):
    """release_server.py imports cleanly with the finder installed (utils.misc, wan.modules.vae, v2v, settings, ...
    from the checkout; utils.wan_wrapper / demo_utils.vae_block3 from this repo) and GenerationSession generates 4
    blocks (window slides at block 2; 'reencode' = the default first-frame re-encode through VAEEncoderWrapper):
    bit-equal with harness/server_loop.py, within bf16 tolerance of the reference-executed golden."""
    res = run("server", mode)
    assert res["blocks"] == 4 and res["latents_rel_l2"] < 1e-2


def test_unmodified_classic_pipeline_on_the_dropins():
    res = run("classic")
    assert res["latents_rel_l2"] < 1e-3
