"""CPU-side checks: the C-ABI library loads and exports every symbol include/plenoxel_abi.h declares, the ctypes
structs have the C layout, host-side geometry matches the oracle, and the product path refuses to run without CUDA."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import plenoxel_oracle as po
from plenoxels_b200 import _lib as L
from plenoxels_b200 import build, ops
from plenoxels_b200.trainer import shard_cameras

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "plenoxel_abi.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(plx_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ("plx_render_fwd", "plx_render_bwd", "plx_adam_step", "plx_generate_rays", "plx_train_step",
              "plx_train_step_host", "plx_gather_nearest", "plx_trilinear_fwd", "plx_composite_fwd"):
        assert s in syms


def test_library_exports_every_declared_symbol(plx_lib):
    raw = C.CDLL(build.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(raw, s), f"{s} declared in plenoxel_abi.h but not exported"
        assert s in L.PROTOTYPES, f"{s} has no ctypes prototype"
    assert plx_lib.plx_version() == 2
    assert plx_lib.plx_num_chunks(600) >= 19 and plx_lib.plx_num_chunks(0) >= 1


def test_ctypes_prototypes_match_the_header_declarations():
    """Every function declared in plenoxel_abi.h has a ctypes prototype with the same number of parameters and, position by
    position, the matching scalar type (pointers and arrays map to c_void_p / POINTER(...))."""
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    decls = re.findall(r"^\s*(?:int|int32_t|const char\s*\*)\s+(plx_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.M | re.S)
    assert len(decls) >= 24 and {d[0] for d in decls} == set(declared_symbols())
    scalars = [("int32_t", C.c_int32), ("uint32_t", C.c_uint32), ("int64_t", C.c_int64), ("uint64_t", C.c_uint64),
               ("float", C.c_float), ("double", C.c_double), ("int", C.c_int)]
    for name, params in decls:
        plist = [" ".join(x.split()) for x in params.split(",")]
        plist = [x for x in plist if x not in ("", "void")]
        _, argtypes = L.PROTOTYPES[name]
        assert len(argtypes) == len(plist), f"{name}: {len(argtypes)} ctypes arguments, {len(plist)} in the header"
        for i, (param, at) in enumerate(zip(plist, argtypes)):
            if "*" in param or "[" in param:
                assert at is L.c_void or at is C.c_char_p or issubclass(at, C._Pointer), f"{name} arg {i}: {param} vs {at}"
                continue
            want = next(ct for cname, ct in scalars if re.match(rf"(const )?{cname}\b", param))
            assert at is want, f"{name} arg {i}: {param} vs {at}"


def test_ctypes_structs_match_c_layout(tmp_path):
    prog = tmp_path / "layout.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "plenoxel_abi.h"\nint main(){'
                    'printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(PlxMarch), sizeof(PlxRays), sizeof(PlxRenderFwd),'
                    'sizeof(PlxRenderBwd), sizeof(PlxTrainStep), offsetof(PlxRenderFwd, loss), offsetof(PlxTrainStep, lr),'
                    'offsetof(PlxTrainStep, loss), sizeof(PlxRenderTrain), offsetof(PlxRenderTrain, beta_over_m),'
                    'sizeof(PlxAdamPeer), offsetof(PlxAdamPeer, result_host), sizeof(PlxPeerSync), offsetof(PlxPeerSync, block_counter),'
                    'offsetof(PlxAdamPeer, sync), offsetof(PlxTrainStep, render_sync));return 0;}')
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(REPO, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [C.sizeof(L.PlxMarch), C.sizeof(L.PlxRays), C.sizeof(L.PlxRenderFwd), C.sizeof(L.PlxRenderBwd),
            C.sizeof(L.PlxTrainStep), L.PlxRenderFwd.loss.offset, L.PlxTrainStep.lr.offset, L.PlxTrainStep.loss.offset,
            C.sizeof(L.PlxRenderTrain), L.PlxRenderTrain.beta_over_m.offset, C.sizeof(L.PlxAdamPeer),
            L.PlxAdamPeer.result_host.offset, C.sizeof(L.PlxPeerSync), L.PlxPeerSync.block_counter.offset,
            L.PlxAdamPeer.sync.offset, L.PlxTrainStep.render_sync.offset]
    assert got == want


def test_every_struct_field_has_the_c_offset(tmp_path):
    """All fields of all ABI structs: offsetof() from the header (compiled with gcc) == the ctypes offset."""
    structs = [L.PlxMarch, L.PlxRays, L.PlxRenderFwd, L.PlxRenderBwd, L.PlxRayGen, L.PlxPeerSync, L.PlxRenderTrain,
               L.PlxAdamPeer, L.PlxTrainStep]
    lines, want = [], []
    for st in structs:
        lines.append(f'printf("%zu\\n", sizeof({st.__name__}));')
        want.append(C.sizeof(st))
        for fname, _ in st._fields_:
            lines.append(f'printf("%zu\\n", offsetof({st.__name__}, {fname}));')
            want.append(getattr(st, fname).offset)
    prog = tmp_path / "offsets.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "plenoxel_abi.h"\nint main(){' + "".join(lines) + "return 0;}")
    exe = tmp_path / "offsets"
    subprocess.run(["gcc", "-I", os.path.join(REPO, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == want


def test_argument_errors_are_reported_not_thrown(plx_lib):
    a = L.PlxRenderFwd()
    assert plx_lib.plx_render_fwd(C.byref(a), None) == -1            # PLX_E_NULL: grid
    assert b"grid" in plx_lib.plx_last_error()
    assert plx_lib.plx_adam_step(None, None, None, None, None, 8, 1e-3, 0.9, 0.999, 1e-8, 1, 1, None) == -1
    assert plx_lib.plx_adam_step(None, None, None, None, None, 0, 1e-3, 0.9, 0.999, 1e-8, 0, 1, None) == -2   # step < 1
    assert plx_lib.plx_generate_rays(None, 1, 4, 4, None, 0.5, None, 4, 2, None, None, None) == -1


def test_tuning_switches_are_the_documented_ones(plx_lib):
    """plx_tune (A/B switches for tools and tests; host-side state only, no launch): every switch DESIGN.md 9 lists exists and
    takes its default back, an unknown name or an out-of-range value is an error code, never an exception."""
    documented = {"adam_skip_same": -1, "adam_blocks_per_sm": 4, "train_wpb": 4, "train_cache_it": 0, "pdl": 1, "packet_tile": 1}
    design = open(os.path.join(REPO, "DESIGN.md")).read()
    for name, default in documented.items():
        assert f"`{name}`" in design, f"{name} is not documented in DESIGN.md"
        assert plx_lib.plx_tune(name.encode(), default) == 0, plx_lib.plx_last_error()
    assert plx_lib.plx_tune(b"no_such_switch", 1) == -3 and b"no_such_switch" in plx_lib.plx_last_error()     # PLX_E_UNSUPPORTED
    assert plx_lib.plx_tune(b"adam_blocks_per_sm", 99) == -2 and plx_lib.plx_tune(b"train_wpb", 3) == -2        # PLX_E_SHAPE
    assert plx_lib.plx_tune(None, 1) == -1
    # the A/B library override is honoured by the loader and by nothing else
    src = open(os.path.join(REPO, "plenoxels_b200", "_lib.py")).read()
    assert src.count("PLX_AB_LIBRARY") == 2


def test_tools_are_importable_python():
    """tools/*.py are measurement drivers that only run on a GPU box; at least they must stay valid Python."""
    import py_compile
    tools = os.path.join(REPO, "tools")
    files = sorted(f for f in os.listdir(tools) if f.endswith(".py"))
    assert files
    for f in files:
        py_compile.compile(os.path.join(tools, f), doraise=True)


@pytest.mark.parametrize("G,pd", [(64, 0.05), (128, 0.025), (256, 0.0125), (93, 0.0125), (7, 0.3)])
def test_grid_origin_matches_oracle(G, pd):
    assert np.array_equal(np.float32(ops.grid_origin((G, G, G), pd)), po.grid_origin((G, G, G), pd))
    assert np.array_equal(np.float32(ops.grid_origin((G, G + 1, G + 2), pd, 3)), po.grid_origin((G, G + 1, G + 2), pd, 3))


def test_cpu_tensors_are_refused(plx_lib):
    grid = torch.zeros(4, 4, 4, 4)
    o, d = torch.zeros(1, 3), torch.ones(2, 3)
    with pytest.raises(L.PlxError, match="no CPU fallback"):
        ops.render_rays(grid, o, d, 8, 0.1, (0, 0, 0), 0.5)
    with pytest.raises(L.PlxError, match="no CPU fallback"):
        ops.composite(torch.zeros(1, 2, 3, 4))
    with pytest.raises(L.PlxError):
        ops.adam_step(grid, grid, grid, grid, grid, 1, 1e-3)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(build, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(L.PlxError, match="no CPU / PyTorch fallback"):
        L.load()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(REPO, "plenoxels_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"


def test_exchange_choice_follows_the_batch_density():
    """pick_exchange (trainer.py): the BASELINE configs take the push exchange (C2: ~1.8 in-bounds samples per cell, C3: 0.03);
    only a batch with many more samples than ~8 per cell falls back to the dense pull exchange."""
    from plenoxels_b200.trainer import pick_exchange
    assert pick_exchange(12800, 600, 128 ** 3) == "push"
    assert pick_exchange(4096, 256, 256 ** 3) == "push"
    assert pick_exchange(1 << 20, 512, 64 ** 3) == "pull"
    assert pick_exchange(0, 600, 128 ** 3) == "push"


@pytest.mark.parametrize("n,world", [(100, 1), (100, 2), (100, 8), (7, 4), (3, 8)])
def test_shard_cameras_partitions(n, world):
    seen = []
    for r in range(world):
        seen += list(shard_cameras(n, r, world))
    assert seen == list(range(n))
    sizes = [len(shard_cameras(n, r, world)) for r in range(world)]
    assert max(sizes) - min(sizes) <= 1


def test_receptive_field_schedule_is_the_references():
    """scripts/train.py:91-92,:108: odd window sizes 93, 91, ..., 3, five steps each, then full resolution (0)."""
    from plenoxels_b200.fit import STEPS_PER_FIELD, receptive_field_at, receptive_field_schedule
    sched = receptive_field_schedule()
    assert sched == list(range(93, 2, -2)) and len(sched) == 46 and STEPS_PER_FIELD == 5
    assert [receptive_field_at(i) for i in (0, 4, 5, 9, 10, 224, 225, 229, 230, 10_000)] == [93, 93, 91, 91, 89, 5, 3, 3, 0, 0]
    assert receptive_field_at(7, schedule=[9, 5]) == 5 and receptive_field_at(10, schedule=[9, 5]) == 0


@pytest.mark.parametrize("n_cells,world", [(2097152, 1), (2097152, 2), (2097152, 8), (1000003, 3), (7, 8), (93 * 91 * 89, 5)])
def test_slab_ranges_partition_the_grid(n_cells, world):
    """Each rank owns a contiguous block of whole cells (multiples of 4 floats); blocks tile [0, 4 * n_cells) in rank order and
    differ by at most one cell — what plx_adam_step_peer's [begin, end) relies on."""
    from plenoxels_b200.trainer import slab_range
    edges = [slab_range(n_cells, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == 4 * n_cells
    for (b0, e0), (b1, e1) in zip(edges, edges[1:]):
        assert e0 == b1
    sizes = [(e - b) // 4 for b, e in edges]
    assert all(b % 4 == 0 and e % 4 == 0 and e >= b for b, e in edges) and max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("n_cells,world", [(2097152, 1), (2097152, 2), (2097152, 8), (16777216, 8), (1000003, 3), (9, 8), (93 * 91 * 89, 5),
                                           (2 ** 31 - 1, 7)])
def test_slab_partition_matches_the_device_owner_function(plx_lib, n_cells, world):
    """Push exchange (PlxPeerGrad): the march sends the gradient of cell `lin` to rank umulhi(lin, owner_mul).  The slabs
    plx_slab_partition hands the optimiser must be exactly the preimages of that function: contiguous, tiling [0, n_cells),
    and the python mirror in trainer.slab_partition must agree with the library."""
    from plenoxels_b200.trainer import slab_partition
    edges = []
    for r in range(world):
        mul, b, e = C.c_uint32(), C.c_int64(), C.c_int64()
        assert plx_lib.plx_slab_partition(n_cells, world, r, C.byref(mul), C.byref(b), C.byref(e)) == 0
        assert (mul.value, b.value, e.value) == slab_partition(n_cells, r, world)
        edges.append((b.value, e.value))
    assert edges[0][0] == 0 and edges[-1][1] == n_cells
    assert all(e0 == b1 for (_, e0), (b1, _) in zip(edges, edges[1:]))
    m = mul.value
    owner = lambda lin: (lin * m) >> 32
    rng = np.random.default_rng(0)
    for r, (b, e) in enumerate(edges):
        probes = [b, e - 1] + list(rng.integers(b, e, size=16)) if e > b else []
        assert all(owner(int(x)) == r for x in probes)
    sizes = [e - b for b, e in edges]
    assert max(sizes) - min(sizes) <= n_cells * n_cells * world // 2 ** 32 + 2     # near-equal: the multiplier is floored to 32 bits


def test_abi_links_and_runs_from_plain_c(plx_lib, tmp_path):
    """The boundary is a C ABI, not a Python one: a C program that includes plenoxel_abi.h and links libplenoxel_b200.so
    must build with gcc and get the documented error codes / messages back (no GPU needed for the argument checks)."""
    src = tmp_path / "abi_smoke.c"
    src.write_text(r"""
#include <stdio.h>
#include <string.h>
#include "plenoxel_abi.h"
int main(void) {
    if (plx_version() != PLX_ABI_VERSION) return 1;
    if (plx_num_chunks(600) < 19) return 2;
    PlxRenderFwd f; memset(&f, 0, sizeof f);
    if (plx_render_fwd(&f, NULL) != -1 || !strstr(plx_last_error(), "grid")) return 3;      /* PLX_E_NULL */
    if (plx_render_fwd(NULL, NULL) != -1) return 4;
    if (plx_adam_step(NULL, NULL, NULL, NULL, NULL, 0, 1e-3, 0.9, 0.999, 1e-8, 0, 1, NULL) != -2) return 5;   /* step < 1 */
    PlxAdamPeer p; memset(&p, 0, sizeof p);
    if (plx_adam_step_peer(&p, NULL) >= 0) return 6;                                         /* world == 0 */
    PlxAdamSlab sl; memset(&sl, 0, sizeof sl);
    if (plx_adam_step_slab(&sl, NULL) >= 0) return 7;                                        /* world == 0 */
    uint32_t mul; int64_t b, e;
    if (plx_slab_partition(2097152, 8, 3, &mul, &b, &e) != 0 || mul != 16384u || b != 3 * 262144 || e != 4 * 262144) return 8;
    if (plx_slab_partition(4, 8, 0, &mul, &b, &e) != PLX_E_SHAPE) return 9;                  /* fewer cells than ranks */
    if (plx_peer_barrier(NULL, 0, 1, 0, 1, NULL, NULL) != PLX_E_NULL) return 10;
    printf("abi ok: %s\n", plx_last_error());
    return 0;
}
""")
    exe = tmp_path / "abi_smoke"
    libdir = os.path.dirname(build.LIB_PATH)
    subprocess.run(["gcc", "-Wall", "-Werror", "-I", os.path.join(REPO, "include"), str(src), "-o", str(exe), "-L", libdir,
                    "-lplenoxel_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "abi ok" in out.stdout
