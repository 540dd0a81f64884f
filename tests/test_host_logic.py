"""CPU-side checks of everything above the C ABI: the library exports what include/epoch_b200.h
declares, the host mirror of EPOCH's decomposition / grid arithmetic agrees with the oracle's
restatement (mpi_routines.F90:317-351, utilities.f90:343-421), the N > 1 set-up path works
across real processes (gloo, world_size 2), and the product path fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from epoch_b200 import deck as D
from epoch_b200 import lib as epb_lib
from epoch_b200 import pic
from oracle.oracle import Oracle
from tests import decks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "epoch_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(epb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _header_functions()
    assert len(names) >= 25
    L = epb_lib.load()
    for n in names:
        assert hasattr(L, n), f"libepoch_b200.so does not export {n}"
    assert sorted(epb_lib.SYMBOLS) == names      # the binding covers the whole header
    assert b"sm_100a" in L.epb_version()


def test_abi_struct_sizes_match_binding():
    L = epb_lib.load()
    info = (C.c_int32 * 4)()
    assert L.epb_abi_info(info) == 0
    assert info[0] == C.sizeof(epb_lib.Config)
    assert info[1] == C.sizeof(epb_lib.SpeciesCfg)
    assert info[2] == D.NG and info[3] == 9


def test_argument_errors_without_gpu():
    L = epb_lib.load()
    assert L.epb_create(None, None, None) == 1            # EPB_ERR_ARG
    assert L.epb_push(None) == 1
    assert L.epb_fields_half(None) == 1
    assert L.epb_last_error(None) == b"null handle"


def test_product_path_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pic.EpbError):
        pic.Simulation(decks.thermal(2, (16, 16), ppc=2))


def test_unsupported_configuration_is_refused():
    """A FIELD boundary code the device path does not implement (here: `thermal`, which efield_bcs / bfield_bcs give
    no edge condition at all, boundary.F90:808-907) returns EPB_ERR_UNSUPPORTED instead of silently diverging.
    (Thermal PARTICLE boundaries are implemented since round 2, tests/test_thermal_bc.py.)"""
    dk = decks.thermal(2, (16, 16), ppc=2, bc="thermal")
    with pytest.raises(pic.EpbError, match="code 3"):
        pic.Simulation(dk)


@pytest.mark.parametrize("ndims,n,nproc", [
    (1, (203,), (3, 1, 1)), (2, (50, 37), (2, 3, 1)), (2, (64, 64), (2, 4, 1)), (3, (21, 17, 19), (2, 2, 2)),
    (3, (16, 23, 9), (1, 3, 1)),
])
@pytest.mark.parametrize("bc", ["periodic", "reflect"])
def test_rank_geometry_matches_oracle(ndims, n, nproc, bc):
    dk = decks.thermal(ndims, n, ppc=1, nproc=nproc, bc=bc)
    o = Oracle(dk)
    assert o.nranks == dk.nranks()
    periods = [d < ndims and bc == "periodic" for d in range(3)]
    lo, hi = o.outer()
    for r in range(o.nranks):
        info, geo = o.rank_info(r), pic.rank_geometry(dk, r)
        assert info["n"] == geo["n"] and info["gmin"] == geo["gmin"]
        assert list(info["coords"]) == list(geo["coords"])
        assert info["is_bnd"] == geo["is_bnd"]
        for key in ("grid_min_local", "min_local", "max_local"):
            assert info[key][:ndims] == geo[key][:ndims], key      # bit-exact doubles
        assert lo[:ndims] == geo["min_outer"][:ndims] and hi[:ndims] == geo["max_outer"][:ndims]
        assert info["neighbour"] == pic._neighbour_table(dk, r, periods)


def test_cell_ranges_remainder_rule():
    # mpi_routines.F90:317-351: the first nxp ranks get nx0 cells, the rest nx0 + 1
    dk = decks.thermal(1, (203,), ppc=1, nproc=(4, 1, 1))
    mins, maxs = dk.cell_ranges(0)
    sizes = [b - a + 1 for a, b in zip(mins, maxs)]
    assert sizes == [50, 51, 51, 51] and mins[0] == 1 and maxs[-1] == 203


def test_bench_split_rule():
    sys.path.insert(0, ROOT)
    import bench
    assert [bench.split_2d(k) for k in (1, 2, 4, 8)] == [(1, 1), (1, 2), (2, 2), (2, 4)]


def test_dump_clock_matches_reference_laser_dump_count():
    # epoch1d/tests/laser: dt_snapshot = 8 fs, t_end = 50 fs -> dumps 0000..0007
    dk = decks.laser1d()
    o = Oracle(dk)
    times = []
    D.run(dk, o, [0], lambda step, t: times.append(t))
    assert len(times) == 8
    assert all(times[k] >= 8e-15 * k for k in range(7))


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from tests import decks
from epoch_b200 import pic
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
dk = decks.thermal(2, (40, 33), ppc=1, nproc=(1, 2, 1))
geo = pic.rank_geometry(dk, rank)
nb = pic._neighbour_table(dk, rank, [True, True, False])
# the N > 1 set-up of bench.py: rank 0 makes the 128-byte communicator id, everyone receives it
idt = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    idt = torch.arange(128, dtype=torch.uint8)
dist.broadcast(idt, 0)
assert idt.tolist() == list(range(128))
tabs = [torch.zeros(27, dtype=torch.int64) for _ in range(2)]
dist.all_gather(tabs, torch.tensor(nb, dtype=torch.int64))
for q in range(27):                      # send direction q of a == receive direction 26-q of b
    b = int(tabs[rank][q])
    if b >= 0:
        assert int(tabs[b][26 - q]) == rank, (rank, q)
ext = [torch.zeros(4, dtype=torch.int64) for _ in range(2)]
dist.all_gather(ext, torch.tensor(geo["n"][:2] + geo["gmin"][:2], dtype=torch.int64))
assert int(ext[0][1] + ext[1][1]) == 33 and int(ext[1][3]) == int(ext[0][1]) + 1
# max-over-ranks timing reduction used by bench.py
t = torch.tensor([1.0 + rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert float(t) == 2.0
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_setup_over_gloo(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT, port=port))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, out) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, out
        assert f"ok {r}" in out
