"""oracle/ -- TEST INFRASTRUCTURE.

CPU restatement of the BMAGWA hot path (oracle.c) plus, when built, the
unmodified reference compiled against shims (oracle/_ref).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; nothing under bmagwa_b200/ does.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = "/root/reference"


def build(ref: bool = True, quiet: bool = True) -> None:
    """Compile liboracle.so and, when /root/reference exists here, oracle/_ref."""
    out = subprocess.DEVNULL if quiet else None
    subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=out)
    if ref and os.path.isdir(os.path.join(REFERENCE_ROOT, "src")):
        subprocess.check_call(["make", "-C", HERE, "ref"], stdout=out)
