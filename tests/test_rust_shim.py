"""rust/dust-cuda-sys/src/lib.rs (the binding a Dust maintainer adds; not compilable here: no Rust toolchain) against
include/dust_b200.h: every entry point declared, with the same number of parameters, and every struct with the same fields in the
same order."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _strip(s):
    return re.sub(r"/\*.*?\*/", "", s, flags=re.S)


def _header():
    return _strip(open(os.path.join(ROOT, "include", "dust_b200.h")).read())


def _rust():
    return re.sub(r"//.*", "", open(os.path.join(ROOT, "rust", "dust-cuda-sys", "src", "lib.rs")).read())


def test_every_entry_point_is_bound_with_the_same_arity():
    h, r = _header(), _rust()
    cfun = {m.group(1): m.group(2) for m in re.finditer(r"\b(dust_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", h)}
    rfun = {m.group(1): m.group(2) for m in re.finditer(r"pub fn (dust_[a-z0-9_]+)\s*\(([^;]*?)\)\s*->", r, flags=re.S)}
    assert set(cfun) == set(rfun), set(cfun) ^ set(rfun)
    for name, args in cfun.items():
        n_c = 0 if args.strip() in ("", "void") else args.count(",") + 1
        n_r = 0 if not rfun[name].strip() else rfun[name].count(",") + 1
        assert n_c == n_r, (name, n_c, n_r)


def test_every_struct_has_the_same_fields_in_the_same_order():
    h, r = _header(), _rust()
    cst = {}
    for m in re.finditer(r"typedef struct (\w+)\s*\{(.*?)\}\s*\1\s*;", h, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            # "float o[3]", "uint32_t *hit_instance, *hit_leaf", "const DustGBuffer *gbuffer", "uint32_t n_levels, _pad"
            first, *rest = decl.split(",")
            names = [re.findall(r"(\w+)\s*(?:\[[^\]]*\])*\s*$", first.strip())[0]] + [re.findall(r"(\w+)\s*(?:\[[^\]]*\])*\s*$", x.strip())[0] for x in rest]
            fields += names
        cst[m.group(1)] = fields
    rst = {m.group(1): re.findall(r"pub (\w+)\s*:", m.group(2)) for m in re.finditer(r"pub struct (\w+)\s*\{\n(.*?)\n\}", r, flags=re.S)}
    assert len(cst) >= 19
    for name, fields in cst.items():
        assert name in rst, name
        assert rst[name] == fields, (name, fields, rst[name])
    opaque = set(re.findall(r"typedef struct (\w+) \1;", h))
    for name in opaque:
        assert re.search(r"pub struct %s \{ _p: \[u8; 0\] \}" % name, r), name


def test_constants_match():
    h, r = _header(), _rust()
    for m in re.finditer(r"\b(DUST_[A-Z0-9_]+)\s*=\s*(\d+)u?", h):
        name, val = m.group(1), int(m.group(2))
        rm = re.search(r"pub const %s: \w+ = (\d+);" % name, r)
        assert rm and int(rm.group(1)) == val, name
    assert re.search(r"#define DUST_ABI_VERSION (\d+)", h).group(1) == re.search(r"pub const DUST_ABI_VERSION: i32 = (\d+);", r).group(1)
