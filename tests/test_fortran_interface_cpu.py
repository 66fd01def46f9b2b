"""The Fortran half of the drop-in boundary (fortran/*.F90) against include/neci_gpu.h.

No Fortran compiler exists in this image, so nothing can compile the ISO_C_BINDING interface block.  What can go
wrong in such a block is mechanical -- a missing symbol, an argument out of order, a scalar passed by reference, the
wrong integer kind, a struct field out of place, a stale enum value -- and all of that is checked here by parsing both
files: every exported prototype of the header must have a `bind(c, name=...)` interface with the same number of
arguments in the same order, each with the C type and the passing convention (VALUE or by reference) the prototype
asks for; `type, bind(c) :: neci_gpu_config` must list the header's fields in order with matching kinds; the
NECI_ST_* / NECI_FLAG_* / NECI_SYS_* parameters must equal the header's values; and every neci_gpu_* call in the shim
(fortran/perform_fcimc_cyc_gpu.F90) must pass as many arguments as its interface declares.  The shim is also checked
for the two duties of PerformFCIMCycPar the first round's sketch forgot (end_iter_stats' SumWalkersCyc/SumWalkersOut
and update_iter_data, src/fcimc_helper.F90:1489-1499, src/fcimc_iter_utilities.F90:1442-1451)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "neci_gpu.h")
IFACE = os.path.join(ROOT, "fortran", "neci_gpu_interface.F90")
SHIM = os.path.join(ROOT, "fortran", "perform_fcimc_cyc_gpu.F90")

C2F = {"int32_t": "integer(c_int32_t)", "int64_t": "integer(c_int64_t)", "uint64_t": "integer(c_int64_t)",
       "double": "real(c_double)", "uint8_t": "integer(c_int8_t)", "int": "integer(c_int)"}


def _strip_c_comments(src):
    return re.sub(r"/\*.*?\*/", " ", src, flags=re.S)


def header_prototypes():
    """name -> (return type, [(ctype, is_pointer, pointer_depth, is_const, argname)])"""
    src = _strip_c_comments(open(HEADER).read())
    protos = {}
    for m in re.finditer(r"(?m)^\s*((?:const\s+)?\w+\s*\**)\s*(neci_gpu_\w+)\s*\(([^;{]*?)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3)
        out = []
        for a in [x.strip() for x in args.replace("\n", " ").split(",") if x.strip()]:
            arr = re.search(r"\[\s*\d*\s*\]$", a)
            if arr:
                a = a[:arr.start()].strip()
            mm = re.match(r"^(const\s+)?(struct\s+)?(\w+)\s*(\*{0,2})\s*(\w+)$", a)
            assert mm, (name, a)
            depth = len(mm.group(4)) + (1 if arr else 0)
            out.append(dict(ctype=mm.group(3), depth=depth, const=bool(mm.group(1)), name=mm.group(5)))
        protos[name] = (ret, out)
    return protos


def _fortran_lines(path):
    """Logical lines: comments stripped, continuation lines joined."""
    out, cur = [], ""
    for raw in open(path).read().splitlines():
        line = raw.split("!")[0].rstrip() if "'" not in raw.split("!")[0] or raw.count("'") % 2 == 0 else raw.rstrip()
        line = line.strip()
        if not line:
            continue
        if line.startswith("&"):
            line = line[1:].strip()
        if line.endswith("&"):
            cur += line[:-1].strip() + " "
            continue
        out.append((cur + line).strip())
        cur = ""
    return out


def fortran_interfaces():
    """bind name -> dict(args=[names], decl={name: (type, value, intent, is_array)}, result=(name, type))"""
    lines = _fortran_lines(IFACE)
    res, cur = {}, None
    for ln in lines:
        low = ln.lower()
        m = re.match(r"function\s+(\w+)\s*\(([^)]*)\)\s*result\s*\((\w+)\)\s*bind\s*\(\s*c\s*,\s*name\s*=\s*'(\w+)'\s*\)", ln, flags=re.I)
        if m:
            cur = dict(fname=m.group(1), args=[a.strip() for a in m.group(2).split(",") if a.strip()], result=m.group(3),
                       bind=m.group(4), decl={})
            continue
        if cur is None:
            continue
        if low.startswith("end function"):
            res[cur["bind"]] = cur
            cur = None
            continue
        if low.startswith("import"):
            continue
        m = re.match(r"^(.*?)::\s*(.*)$", ln)
        if m:
            spec, names = m.group(1), m.group(2)
            parts = [p.strip().lower() for p in re.split(r",(?![^()]*\))", spec) if p.strip()]
            ftype = parts[0].replace(" ", "")
            value = "value" in parts
            intent = next((p for p in parts if p.startswith("intent")), None)
            for nm in re.split(r",(?![^()]*\))", names):
                nm = nm.strip()
                is_arr = "(" in nm
                cur["decl"][nm.split("(")[0].strip()] = (ftype, value, intent, is_arr)
    return res


def test_every_exported_symbol_is_bound_with_matching_arguments():
    protos = header_prototypes()
    ifs = fortran_interfaces()
    lib_path = os.path.join(ROOT, "neci_stable_b200", "libneci_gpu.so")
    assert len(protos) >= 32
    assert set(protos) == set(ifs), (sorted(set(protos) - set(ifs)), sorted(set(ifs) - set(protos)))
    if os.path.exists(lib_path):                                   # the built library exports exactly these
        lib = ctypes.CDLL(lib_path)
        for name in protos:
            assert hasattr(lib, name), name
    for name, (ret, cargs) in protos.items():
        f = ifs[name]
        assert f["fname"] == name, "Fortran name differs from the bind name: %s" % name
        assert len(f["args"]) == len(cargs), (name, f["args"], [a["name"] for a in cargs])
        # result kind
        rtype = f["decl"][f["result"]][0]
        want_ret = {"int": "integer(c_int)", "int64_t": "integer(c_int64_t)"}.get(ret.replace("const", "").strip(), None)
        if "*" in ret:
            want_ret = "type(c_ptr)"
        assert rtype == want_ret, (name, rtype, ret)
        for fa, ca in zip(f["args"], cargs):
            assert fa in f["decl"], (name, fa, "argument not declared")
            ftype, value, intent, is_arr = f["decl"][fa]
            ct, depth = ca["ctype"], ca["depth"]
            where = "%s(%s <- %s)" % (name, fa, ca["name"])
            if depth == 0:                                         # C scalar by value
                assert value and not is_arr, where + ": scalar must have the VALUE attribute"
                assert ftype == C2F[ct], (where, ftype, ct)
            elif ct in ("neci_gpu_engine", "void"):
                assert ftype == "type(c_ptr)", where
                if depth == 1:
                    assert value, where + ": opaque handle is passed by value"
                else:
                    assert not value and intent == "intent(out)", where + ": handle** is an out argument by reference"
            elif ct == "neci_gpu_config":
                assert ftype == "type(neci_gpu_config)" and not value and intent == "intent(in)", where
            else:                                                  # typed pointer
                assert depth == 1, where
                if value:                                          # nullable pointer: type(c_ptr), value
                    assert ftype == "type(c_ptr)", where
                else:
                    assert ftype == C2F[ct], (where, ftype, ct)
                    assert intent is not None, where
                    if ca["const"]:
                        assert intent == "intent(in)", (where, intent)
                    else:
                        assert intent in ("intent(out)", "intent(inout)"), (where, intent)


def test_config_struct_fields_match_the_header_in_order_and_kind():
    src = _strip_c_comments(open(HEADER).read())
    body = re.search(r"typedef struct neci_gpu_config \{(.*?)\} neci_gpu_config;", src, flags=re.S).group(1)
    cfields = []
    for stmt in [x.strip() for x in body.split(";") if x.strip()]:
        m = re.match(r"^(const\s+)?(\w+)\s*(\*?)\s*(.*)$", stmt)
        ctype, ptr = m.group(2), m.group(3)
        for nm in [x.strip() for x in m.group(4).split(",")]:
            star = ptr or ("*" if nm.startswith("*") else "")
            cfields.append((nm.lstrip("*").strip(), "type(c_ptr)" if star else C2F[ctype]))
    lines = _fortran_lines(IFACE)
    i0 = next(i for i, l in enumerate(lines) if re.match(r"type\s*,\s*bind\(c\)\s*::\s*neci_gpu_config", l, flags=re.I))
    i1 = next(i for i, l in enumerate(lines) if i > i0 and l.lower().startswith("end type"))
    ffields = []
    for l in lines[i0 + 1:i1]:
        spec, names = l.split("::")
        for nm in names.split(","):
            ffields.append((nm.strip(), spec.strip().lower().replace(" ", "")))
    assert ffields == cfields
    # and the Python binding's ctypes structure has the same size as a C compiler's layout of these fields
    from neci_stable_b200 import capi
    size = 0
    for _, ft in cfields:
        w = 4 if ft in ("integer(c_int32_t)",) else 8
        size = (size + w - 1) // w * w + w
    assert ctypes.sizeof(capi.Config) == (size + 7) // 8 * 8


def test_constants_match_the_header():
    src = _strip_c_comments(open(HEADER).read())
    want = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(NECI_(?:FLAG|SYS)_\w+)\s+(\d+)", src)}
    body = re.search(r"enum neci_stat_index \{(.*?)\};", src, flags=re.S).group(1)
    val = 0
    for item in [x.strip() for x in body.split(",") if x.strip()]:
        if "=" in item:
            nm, v = [t.strip() for t in item.split("=")]
            val = int(v)
        else:
            nm = item
        want[nm] = val
        val += 1
    got = {}
    for l in _fortran_lines(IFACE):
        m = re.match(r"integer\(c_int\)\s*,\s*parameter\s*::\s*(NECI_\w+)\s*=\s*(\d+)", l, flags=re.I)
        if m:
            got[m.group(1)] = int(m.group(2))
    assert got == want
    from neci_stable_b200 import capi
    assert capi.ST_COUNT == want["NECI_ST_COUNT"]
    for k, v in capi.ST.items():
        assert want["NECI_ST_" + k] == v


def _call_args(text, start):
    depth, i, args, cur = 0, start, [], ""
    while True:
        ch = text[i]
        if ch == "(":
            depth += 1
            if depth > 1:
                cur += ch
        elif ch == ")":
            depth -= 1
            if depth == 0:
                args.append(cur.strip())
                return [a for a in args if a]
            cur += ch
        elif ch == "," and depth == 1:
            args.append(cur.strip()); cur = ""
        else:
            cur += ch
        i += 1


def test_shim_calls_match_the_interfaces_and_do_what_perform_fcimc_cyc_par_does():
    ifs = fortran_interfaces()
    text = "\n".join(_fortran_lines(SHIM))
    calls = list(re.finditer(r"\b(neci_gpu_\w+)\s*\(", text))
    names = set()
    for m in calls:
        name = m.group(1)
        if name in ("neci_gpu_config", "neci_gpu_interface"):
            continue
        assert name in ifs, name
        n = len(_call_args(text, m.end() - 1))
        assert n == len(ifs[name]["args"]), (name, n, ifs[name]["args"])
        names.add(name)
    assert {"neci_gpu_init", "neci_gpu_iterate", "neci_gpu_upload_walkers", "neci_gpu_download_walkers",
            "neci_gpu_finalize", "neci_gpu_last_error"} <= names
    low = text.lower()
    # end_iter_stats: before the engine call, on the TotParts the walker loop saw
    a, b = low.index("sumwalkerscyc = sumwalkerscyc + totparts"), low.index("neci_gpu_iterate(")
    assert a < b and "sumwalkersout = sumwalkersout + totparts" in low
    # update_iter_data
    assert "iter_data%update_growth = iter_data%update_growth + iter_data%nborn" in low
    assert "iter_data%update_iters = iter_data%update_iters + 1" in low
    # every statistic named in the shim exists in the interface module
    consts = {m.group(1) for m in re.finditer(r"\b(NECI_ST_\w+)\b", text)}
    declared = {m.group(1) for l in _fortran_lines(IFACE) for m in [re.match(r".*::\s*(NECI_ST_\w+)\s*=", l)] if m}
    assert consts <= declared, consts - declared
