"""CPU: the host side of the resident-mode protocol (csrc/mppi_resident_host.h — the code the library runs) against an
emulated grid: a host thread that follows the device side of the protocol as csrc/mppi_resident.cuh implements it.

What it pins without a GPU: record packing for f32 / f64 states and 64-bit Philox counters, self-validating words, the
idle-exit race (a command posted while the grid is leaving is neither lost nor run twice: the host relaunches and the
record is still in the box), stop records consuming a sequence number, a fresh controller on a box full of old words,
reseeding, `sync` waiting for the finisher's done word, argument errors."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_resident_protocol_against_emulated_grid(tmp_path):
    exe = str(tmp_path / "resident_protocol_harness")
    src = os.path.join(ROOT, "tests", "resident_protocol_harness.cpp")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, src], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ALL OK" in r.stdout
