"""The Rust side of the boundary ships as source (no rustc / cargo in this image): what CAN be checked is checked.

  * rust/diee-sys/src/lib.rs is exactly what tools/gen_diee_sys.py generates from include/diee.h today;
  * every `repr(C)` struct in it has the size and field offsets the C compiler gives the header's struct (computed here
    with Rust's repr(C) layout rules from the parsed Rust source, against a C program printing sizeof / offsetof);
  * every function the header declares is bound, with the same number of parameters, and every constant agrees;
  * the safe crate only calls functions the -sys crate exports, and implements every method of the reference's
    `LearnableGame` trait (src/base.rs:8-51) for both games."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tools"))
SYS_RS = os.path.join(ROOT, "rust", "diee-sys", "src", "lib.rs")
SAFE = os.path.join(ROOT, "rust", "diee", "src")

PRIM = {"i8": (1, 1), "u8": (1, 1), "i16": (2, 2), "u16": (2, 2), "i32": (4, 4), "u32": (4, 4), "f32": (4, 4),
        "i64": (8, 8), "u64": (8, 8), "f64": (8, 8)}


def _rust_structs():
    src = open(SYS_RS).read()
    out = {}
    for name, body in re.findall(r"#\[repr\(C\)\]\s*#\[derive\([^)]*\)\]\s*pub struct (\w+) \{(.*?)\n\}", src, flags=re.S):
        out[name] = [(f, t.strip()) for f, t in re.findall(r"pub (\w+): ([^,]+),", body)]
    return out


def _layout(structs, name):
    """(size, align, {field: offset}) by the repr(C) rules: fields in order, each aligned to its own alignment, the
    struct padded to its largest alignment"""
    off, align, offs = 0, 1, {}
    for f, t in structs[name]:
        m = re.match(r"\[(\w+); (\d+)\]", t)
        base, count = (m.group(1), int(m.group(2))) if m else (t, 1)
        if base in PRIM:
            sz, al = PRIM[base]
        else:
            sz, al, _ = _layout(structs, base)
        off = (off + al - 1) // al * al
        offs[f] = off
        off += sz * count
        align = max(align, al)
    return (off + align - 1) // align * align, align, offs


def test_sys_crate_is_what_the_generator_makes_from_the_header():
    import gen_diee_sys
    assert open(SYS_RS).read() == gen_diee_sys.generate(), "run `python tools/gen_diee_sys.py` after editing include/diee.h"


def test_struct_layouts_match_the_c_compiler(tmp_path):
    structs = _rust_structs()
    assert len(structs) >= 10
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "diee.h"', 'int main(void) {']
    for name, fields in structs.items():
        lines.append(f'  printf("{name} size %zu\\n", sizeof({name}));')
        for f, _ in fields:
            lines.append(f'  printf("{name} {f} %zu\\n", offsetof({name}, {f}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = {}
    for ln in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines():
        s, f, v = ln.split()
        got[(s, f)] = int(v)
    for name in structs:
        size, _, offs = _layout(structs, name)
        assert got[(name, "size")] == size, (name, got[(name, "size")], size)
        for f, o in offs.items():
            assert got[(name, f)] == o, (name, f)
    # and the numpy dtypes the Python host uses are the same layouts
    from die_e_b200 import _ffi
    for name, dt in (("diee_bg_state", _ffi.BG_STATE), ("diee_move", _ffi.MOVE), ("diee_ttt_state", _ffi.TTT_STATE),
                     ("diee_mcts_cfg", _ffi.MCTS_CFG), ("diee_node", _ffi.NODE), ("diee_search_stats", _ffi.SEARCH_STATS),
                     ("diee_anode", _ffi.ANODE), ("diee_traj_record", _ffi.TRAJ), ("diee_selfplay_opts", _ffi.SELFPLAY_OPTS),
                     ("diee_selfplay_report", _ffi.SELFPLAY_REPORT)):
        assert dt.itemsize == got[(name, "size")], name
        for f in dt.names:
            assert dt.fields[f][1] == got[(name, f)], (name, f)


def test_every_header_function_and_constant_is_bound():
    import gen_diee_sys
    defines, enums, structs, opaque, funcs = gen_diee_sys.parse(open(gen_diee_sys.HEADER).read())
    rs = open(SYS_RS).read()
    from die_e_b200 import _ffi
    assert sorted(n for _, n, _ in funcs) == sorted(_ffi.SYMBOLS)          # the header, the ctypes host and the crate agree
    for _, name, args in funcs:
        m = re.search(r"pub fn %s\((.*?)\)" % name, rs)
        assert m, name
        n_c = 0 if args.strip() in ("", "void") else len(args.split(","))
        n_rs = 0 if not m.group(1).strip() else len(m.group(1).split(","))
        assert n_c == n_rs, name
    for name, val in enums:
        assert re.search(r"pub const %s: i32 = %d;" % (name, val), rs), name
    for name, val in defines:
        if name == "DIEE_H":
            continue
        num = re.sub(r"[()u]", "", val.strip())
        assert re.search(r"pub const %s: \w+ = %s;" % (name, re.escape(num)), rs), name


def test_safe_crate_uses_only_exported_functions_and_implements_the_trait():
    rs = open(SYS_RS).read()
    exported = set(re.findall(r"pub fn (diee_\w+)\(", rs))
    consts = set(re.findall(r"pub const (DIEE_\w+):", rs))
    structs = set(re.findall(r"pub struct (diee_\w+)", rs))
    used_fn, used_other = set(), set()
    for fn in os.listdir(SAFE):
        src = open(os.path.join(SAFE, fn)).read()
        used_fn |= set(re.findall(r"sys::(diee_\w+)\(", src))
        used_other |= set(re.findall(r"sys::(DIEE_\w+|diee_\w+)\b(?!\()", src))
    assert used_fn and used_fn <= exported, used_fn - exported
    assert used_other <= consts | structs | exported, used_other - consts - structs - exported
    # every entry point of the path is reached from the safe crate
    for f in ("diee_mcts_search", "diee_alpha_search", "diee_selfplay_run", "diee_net_create", "diee_net_forward",
              "diee_bg_valid_moves", "diee_bg_apply_moves", "diee_bg_encode_moves", "diee_bg_decode_moves",
              "diee_bg_encode_states", "diee_comm_init"):
        assert f in used_fn, f
    trait = open(os.path.join(SAFE, "base.rs")).read()
    methods = set(re.findall(r"\n    fn (\w+)\(", trait)) - {"as_tch_tensor"}
    ref_methods = {"new", "name", "get_valid_moves", "apply_move", "roll_die", "skip_turn", "get_player", "check_winner",
                   "as_tensor", "decode", "encode", "get_id", "set_id", "to_pretty_str"}          # src/base.rs:26-50
    assert methods == ref_methods, methods ^ ref_methods
    for game in ("backgammon.rs", "tictactoe.rs"):
        src = open(os.path.join(SAFE, game)).read()
        impl = src[src.index("impl LearnableGame for"):]
        have = set(re.findall(r"\n    fn (\w+)\(", impl))
        need = ref_methods - ({"roll_die"} if game == "tictactoe.rs" else set())
        assert need <= have, (game, need - have)
        for c in ("EMPTY_MOVE", "IS_DETERMINISTIC", "ACTION_SPACE_SIZE", "N_INPUT_CHANNELS", "CONV_OUTPUT_SIZE", "N_FILTERS", "N_RES_BLOCKS"):
            assert re.search(r"const %s:" % c, impl), (game, c)
