"""CPU checks of bench.py's bookkeeping: DRAM-traffic figures are only reported while the kernel source they were captured
from is unchanged, and the strong-scaling share of the batch is what BASELINE config[2] says."""
import hashlib
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    argv, sys.argv = sys.argv, sys.argv[:1]
    try:
        return importlib.import_module("bench")
    finally:
        sys.argv = argv


def test_traffic_entries_are_current_or_flagged(tmp_path, monkeypatch):
    bench = _bench()
    entries = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["entries"]
    assert entries, "profiles/traffic.json has no captures"
    for e in entries:
        val, src = bench.load_traffic(e["kernel"], e["config"])
        fresh = all(hashlib.sha1(open(os.path.join(ROOT, rel), "rb").read()).hexdigest() == sha for rel, sha in e["src_sha1"].items())
        if fresh:
            assert val == float(e["dram_bytes"]) and src == e["source"]
        else:
            assert val is None and "stale" in src
    assert bench.load_traffic("no_such_kernel", "x") == (None, "no capture for this kernel / configuration")
    # a capture whose source has changed since must not be reported
    fake = {"entries": [dict(entries[0], src_sha1={k: "0" * 40 for k in entries[0]["src_sha1"]})]}
    (tmp_path / "profiles").mkdir()
    (tmp_path / "profiles" / "traffic.json").write_text(json.dumps(fake))
    for rel in entries[0]["src_sha1"]:
        dst = tmp_path / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        dst.write_bytes(open(os.path.join(ROOT, rel), "rb").read())
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    val, why = bench.load_traffic(entries[0]["kernel"], entries[0]["config"])
    assert val is None and why.startswith("stale")


def test_numa_binding_never_raises():
    bench = _bench()

    class _Props(object):
        pci_bus_id, pci_device_id, pci_domain_id = 255, 31, 65535        # no such device: the helper must shrug

    class _Cuda(object):
        @staticmethod
        def get_device_properties(i):
            return _Props()

    class _Torch(object):
        cuda = _Cuda()

    before = os.sched_getaffinity(0)
    info = bench.bind_to_gpu_numa(_Torch(), 0, 8)
    assert set(info) >= {"numa_node", "cpus"} and info["numa_node"] is None
    assert os.sched_getaffinity(0) == before
