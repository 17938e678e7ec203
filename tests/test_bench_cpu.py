"""Host-side pieces of bench.py that need no GPU: the roofline's `traffic` may only quote an ncu capture of the SAME kernel source."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_traffic_is_only_quoted_for_the_captured_source_version():
    table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    assert "blend_bwd2_kernel" in table
    for name, e in table.items():
        assert {"workload", "source_sha16", "dram_bytes", "time_us", "issue_active_pct", "capture", "commit"} <= set(e), name
    e = table["blend_bwd2_kernel"]
    src = os.path.join(ROOT, "street_gaussians_b200", "csrc", "blend_bwd2.cu")
    current = hashlib.sha256(open(src, "rb").read()).hexdigest()[:16]
    traffic, issue, stale = bench.ncu_traffic("blend_bwd2_kernel", "blend_bwd2.cu", e["workload"])
    if e["source_sha16"] == current:
        assert traffic == e["dram_bytes"] and stale is None
    else:  # an older build was captured: nothing is claimed for this one, the old capture is passed on labelled as such
        assert traffic is None and issue is None
        assert stale["dram_bytes"] == e["dram_bytes"] and "EARLIER build" in stale["note"]
    # another workload, or a kernel without a capture: nothing
    assert bench.ncu_traffic("blend_bwd2_kernel", "blend_bwd2.cu", "not-" + e["workload"])[0] is None
    assert bench.ncu_traffic("no_such_kernel", "blend_bwd2.cu", "C") == (None, None, None)
