"""CPU tests of the boundary and the host logic (no compute calls without a GPU):
 * libace_b200.so loads and exports every symbol include/ace_b200.h and
   include/rt_ant/rt_ant.h declare;
 * creating a context without a GPU fails loudly (no CPU fallback);
 * the multi-rank work split + max-over-ranks timing reduction (gloo, world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, pattern):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(pattern, text)))


def test_c_abi_exports_every_declared_symbol():
    import ace_compiler_b200 as ace
    lib = ace.load_library()
    names = _declared("ace_b200.h", r"\b(ace_[a-z0-9_]+)\s*\(")
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), n
        assert n in ace.SIGNATURES, "python binding missing for " + n


def test_rt_ant_surface_is_exported():
    import ace_compiler_b200 as ace
    lib = ace.load_library()
    text = open(os.path.join(ROOT, "include", "rt_ant", "rt_ant.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"static inline[^{]*\{[^}]*\}", "", text, flags=re.S)  # header inlines
    names = set(re.findall(r"\b([A-Z][A-Za-z0-9_]*)\s*\(", text))
    callbacks = {"Get_context_params", "Get_rt_data_info", "Get_input_count", "Get_output_count",
                 "Get_encode_scheme", "Get_decode_scheme", "Main_graph"}
    macros = {n for n in names if n.isupper()}
    expect = names - callbacks - macros
    # every symbol the emitted ResNets call must be among them
    for n in ["Get_input_data", "Set_output_data", "Degree", "Q_modulus", "P_modulus",
              "Alloc_poly", "Free_poly", "Free_poly_data", "Free_ciph_poly", "Zero_ciph",
              "Copy_ciph", "Level", "Sc_degree", "Num_decomp", "Set_coeffs",
              "Init_ciph_same_scale", "Init_ciph_same_scale_plain", "Init_ciph_same_scale_ciph3",
              "Init_ciph_up_scale_plain", "Init_ciph3_up_scale", "Init_ciph_down_scale",
              "Hw_modadd", "Hw_modmul", "Hw_rotate", "Decomp_modup", "Mod_down", "Rescale",
              "Swk", "Pk0_at", "Pk1_at", "Auto_order", "Bootstrap", "Pt_from_msg",
              "Encode_plain_from_float", "Encode_plain_from_double", "Tm_start", "Tm_taken",
              "Prepare_context", "Finalize_context", "Prepare_input", "Handle_output",
              "Run_main_graph", "Alloc_tensor", "Free_tensor", "Decomp", "Mod_up",
              "Init_ciph_up_scale", "Add_ciph", "Mul_ciph", "Relin", "Rotate_ciph",
              "Rescale_ciph", "Encrypt", "Print_cipher_msg"]:
        assert n in expect, n
    for n in sorted(expect):
        assert hasattr(lib, n), "libace_b200.so does not export " + n


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import ace_compiler_b200 as ace
    with pytest.raises(ace.AceError, match="no CUDA device"):
        ace.Context(1024, 5, 60, 56, 2)


def test_product_never_imports_oracle():
    """the product path must not load, link or import anything under oracle/"""
    for base, _, files in os.walk(os.path.join(ROOT, "ace_compiler_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(base, f)).read()
                for needle in ("liboracle", "libace_ref", "oracle_bindings", "ckks_oracle",
                               "oracle/"):
                    assert needle not in src, (f, needle)


def test_shard_units():
    from ace_compiler_b200.sharding import shard_units
    for total in (0, 1, 7, 8, 33):
        for world in (1, 2, 4, 8):
            got = sorted(sum((shard_units(total, r, world) for r in range(world)), []))
            assert got == list(range(total))
            sizes = [len(shard_units(total, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_units(4, 2, 2)


_WORKER = r'''
import os, sys
sys.path.insert(0, %r)
import torch.distributed as dist
from ace_compiler_b200.sharding import shard_units, max_over_ranks
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
mine = shard_units(10, rank, world)
gathered = [None] * world
dist.all_gather_object(gathered, mine)
assert sorted(sum(gathered, [])) == list(range(10)), gathered
t = max_over_ranks(1.0 + rank, dist)      # rank 1 is the slow one
assert t == float(world), t
dist.barrier()
if rank == 0:
    print("GLOO_OK", t)
dist.destroy_process_group()
'''


def test_two_rank_gloo_split_and_timing(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
                        "29533", str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "GLOO_OK 2.0" in r.stdout
