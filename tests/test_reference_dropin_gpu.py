"""GPU: the UNMODIFIED reference's `YOLO(...).predict()` (JDE task) under `sarpost.patch()` — needs the reference
install `baseline/_ref` (git-ignored, travels with the gpurun snapshot); skipped when it is absent."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_real_predict_under_all_patch_modes(cuda):
    """tools/dropin_predict_check.py in a fresh process (the reference's full package import must not meet the light
    shim other tests load): reference as is, patch(), patch(fused=True), patch(fused=True, defer_state=True); every patched
    run is checked against the CPU oracle on the logits captured in that run (mismatch budget 1e-4)."""
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("no reference install (baseline/_ref) on this machine")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dropin_predict_check.py"), "--device", "0"],
                         capture_output=True, text=True, timeout=900)
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert lines, res.stdout[-1500:] + res.stderr[-3000:]
    rep = json.loads(lines[-1])
    assert res.returncode == 0 and rep.get("ok"), json.dumps(rep, indent=1) + res.stderr[-2000:]
    for name in ("patch", "patch_fused", "patch_fused_split", "patch_fused_split_nchw_emb", "patch_fused_predictor", "patch_fused_defer_state",
                 "patch_fused_defer_state_predictor"):
        assert rep["runs"][name]["ok"], rep["runs"][name]
        assert rep["runs"][name]["box_columns"] == 7 and sum(rep["runs"][name]["detections"]) > 0
    assert rep["runs"]["patch_fused_defer_state"]["levels_channels"] == 64 + 1 + 256  # the state MLP really was skipped in the forward
    assert rep["runs"]["patch_fused_split"]["head_returned_split_levels"] and not rep["runs"]["patch_fused"]["head_returned_split_levels"]
    assert rep["validator_match"]["ok"], rep["validator_match"]
