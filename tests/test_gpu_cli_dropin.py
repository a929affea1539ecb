"""The reference's CLI, built from its UNCHANGED main.c against this library (make -C voxtral.c_b200 voxtral, done by
__graft_entry__.build() in the build container; the binary travels in voxtral.c_b200/_build/), run as a user would run it."""
import os
import re
import subprocess

import pytest

from conftest import ROOT, golden, synth_wav

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "voxtral.c_b200", "_build", "voxtral")


@pytest.mark.skipif(not os.path.exists(CLI), reason="CLI not built (needs /root/reference/main.c at build time)")
def test_reference_cli_runs_on_the_engine(model_dir):
    g = golden("synth_s2_oneshot")
    r = subprocess.run([CLI, "-d", model_dir, "-i", synth_wav(2)], capture_output=True, timeout=600)
    err = r.stderr.decode(errors="replace")
    assert r.returncode == 0, err[-1500:]
    assert r.stdout.rstrip(b"\n") == g["text"].tobytes()          # the reference CLI prints exactly this for the same input
    # the statistics lines benchmark.py parses (voxtral.c:1308-1316, main.c:390)
    assert re.search(r"Audio: 32000 samples \(2\.0 seconds\)", err)
    assert re.search(r"Encoder: 592 mel -> 74 tokens \(\d+ ms\)", err)
    assert re.search(r"Decoder: 36 text tokens \(36 steps\) in \d+ ms \(prefill \d+ ms \+ [\d.]+ ms/step\)", err)
