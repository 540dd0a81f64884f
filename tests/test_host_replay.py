"""host/replay.cpp: the compiled host on the reference's side of the C ABI.  It keeps the particles in EPOCH-style
linked lists (one heap node per particle), and does b200_attach / the PIC loop / b200_download exactly as
fortran/epoch_b200_mod.F90 would (which cannot be compiled here: no Fortran in the image).  The GPU test runs it
as a separate process on a state file and checks what comes back out of its lists against the CPU oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

from tests import decks
from tests.gpu_util import FIELDS, rel_l2, sorted_rows

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPLAY = os.path.join(ROOT, "host", "replay")


def test_replay_builds_and_links_against_the_c_abi():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "host"), "-s"])
    assert os.path.exists(REPLAY)
    r = subprocess.run([REPLAY], capture_output=True, text=True)
    assert r.returncode == 1 and "usage: replay" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("ndims,n", [(2, (48, 32)), (1, (96,)), (3, (12, 10, 9))])
def test_replay_matches_oracle(tmp_path, ndims, n):
    from epoch_b200.pic import Simulation
    from oracle.oracle import Oracle
    dk = decks.thermal(ndims, n, ppc=6, temp_k=2.0e8, two_species=True)
    nsteps = 7
    o = Oracle(dk)
    o.auto_load()
    parts = [o.get_particles(0, isp).copy() for isp in range(2)]
    sim = Simulation(dk, strict_fp=True, sort_interval=2, capacity_factor=2.0)   # for its config structs only
    state, result = str(tmp_path / "state.bin"), str(tmp_path / "result.bin")
    sim.write_replay_state(state, nsteps, {}, parts)
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "host"), "-s"])
    r = subprocess.run([REPLAY, state, result], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["particles"] == sum(p.shape[0] for p in parts) and info["gpu_launches"] > 0
    fields, got = sim.read_replay_result(result)
    sim.close()
    # the oracle does what PROGRAM pic does for these decks (no lasers): init, then the four calls per step
    o.init()
    for _ in range(nsteps):
        o.fields_half(); o.push(); o.current_finish(); o.fields_final()
    for name in FIELDS:
        assert rel_l2(fields[name], o.field(0, name)) <= 1e-12, name
    for isp in range(2):
        a, b = sorted_rows(got[isp]), sorted_rows(o.get_particles(0, isp))
        assert a.shape == b.shape
        assert np.all(np.abs(a - b) <= 1e-9 * np.max(np.abs(b), axis=0))
