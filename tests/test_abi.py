"""The C-ABI library must load (no GPU needed) and export exactly the symbols include/fullbatch_b200.h declares; the
ctypes binding must cover all of them.  No compute call is made here."""
import ctypes
import os
import re
import subprocess

import pytest

from fullbatchtraining_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "fullbatch_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(fb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = header_symbols()
    assert len(names) >= 20
    handle = lib.load()
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"
    assert sorted(lib.EXPORTS) == names, "ctypes binding and header disagree"
    assert handle.fb_version() >= 100


def test_struct_sizes_match_the_header_layout(tmp_path):
    """sizeof / last-member offset of every argument struct as gcc lays the header out == the ctypes mirror"""
    structs = {"fb_tap": (lib.Tap, "b_k0"), "fb_wgrad_tap": (lib.WgradTap, "k_index"),
               "fb_conv_gemm_args": (lib.ConvGemmArgs, "sched_k_iters"), "fb_wgrad_args": (lib.WgradArgs, "ng"),
               "fb_reduce_entry": (lib.ReduceEntry, "vec"), "fb_wprep_entry": (lib.WprepEntry, "wd_gstride"),
               "fb_bn_apply_args": (lib.BnApplyArgs, "reverse"), "fb_bn_bwd_args": (lib.BnBwdArgs, "reverse"),
               "fb_bn_ema_entry": (lib.BnEmaEntry, "c_start")}
    src = tmp_path / "sizes.c"
    body = "".join(f'  printf("{n} %zu %zu\\n", sizeof({n}), offsetof({n}, {last}));\n' for n, (_, last) in structs.items())
    src.write_text(f'#include <stdio.h>\n#include <stddef.h>\n#include "{ROOT}/include/fullbatch_b200.h"\n'
                   f"int main(void) {{\n{body}  return 0;\n}}\n")
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = 0
    for line in out:
        if not line.strip():
            continue
        name, size, off = line.split()
        cls, last = structs[name]
        assert ctypes.sizeof(cls) == int(size), name
        assert getattr(cls, last).offset == int(off), name
        seen += 1
    assert seen == len(structs)


def test_tile_schedule_is_a_host_function_of_one_groups_problem():
    """fb_conv_stats_rows / fb_conv_pair_ok are pure host functions (no GPU needed): rows per group within
    1 .. min(tiles, 148 / N tiles); without statistics (sched_k_iters = 0) or for a single-group policy one row per CTA
    of a wave with a preference for divisors; with statistics the makespan model of DESIGN.md 5.1 (values pinned for
    the ResNet-18 / ResNet-152 shapes the round-2 measurements quote); pairs need an even divisor."""
    L = lib.load(build_if_missing=False)
    rows = L.fb_conv_stats_rows
    for mtg in (1, 2, 7, 16, 64, 100, 256, 1024):
        for n_tiles in (1, 2, 4, 8):
            for pg in (1, 8, 16):
                for k in (0, 1, 4, 9, 72):
                    r = rows(mtg, n_tiles, pg, k)
                    assert 1 <= r <= min(mtg, max(148 // n_tiles, 1))
                    assert r == rows(mtg, n_tiles, pg, k)
                    assert L.fb_conv_pair_ok(mtg, n_tiles, pg, k) in (0, 1)
                    if L.fb_conv_pair_ok(mtg, n_tiles, pg, k):
                        assert r % 2 == 0 and mtg % r == 0
    assert rows(1024, 1, 8, 0) == 128 and rows(1024, 1, 1, 9) == 128   # one row per CTA of a wave, dividing the tiles
    assert rows(1024, 1, 8, 9) == 37 and rows(256, 1, 8, 18) == 37     # ResNet-18 32x32 / 16x16 forward
    assert rows(16, 2, 8, 72) == 8                                     # 4x4x512 forward (CTA pairs: even divisor)
    assert rows(256, 1, 8, 1) == 18 and rows(16, 8, 8, 4) == 2         # ResNet-152 1x1 convolutions


def test_errors_are_reported_not_thrown():
    handle = lib.load()
    rc = handle.fb_flat_scale(None, 0, 1.0, None)  # argument validation happens before any CUDA call
    assert rc == 1001
    assert "fb_flat_scale" in lib.last_error()
    with pytest.raises(RuntimeError):
        lib.check(rc, "fb_flat_scale")
