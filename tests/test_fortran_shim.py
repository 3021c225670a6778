"""The ISO_C_BINDING interfaces of sosie_b200/fortran/sosie_gpu_c.f90 against the prototypes of include/sosie_b200.h.

No Fortran compiler exists in this image, so the shim cannot be built here; what CAN be checked without one is that every
interface binds an entry the header declares, with the same number of arguments, each passed the way the C side expects it
(by value: int32 / size_t / float / double / pointer; by reference: pointer) and with the matching element type."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _c_prototypes():
    src = open(os.path.join(ROOT, "include", "sosie_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    protos = {}
    for m in re.finditer(r"\b([A-Za-z_][\w\s\*]*?)\b(sosie_\w+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        if "typedef" in ret or "(*" in args and name.endswith("_fn"):
            continue
        out = []
        for a in [x.strip() for x in _split_args(args)]:
            if a in ("", "void"):
                continue
            out.append(_c_class(a))
        protos[name] = out
    return protos


def _split_args(s):
    depth, cur, out = 0, "", []
    for ch in s:
        if ch == "(":
            depth += 1
        if ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur); cur = ""
        else:
            cur += ch
    out.append(cur)
    return out


def _c_class(a):
    """(how, element type) of one C parameter"""
    if "*" in a or "_fn" in a:
        base = a.replace("const", "").split("*")[0].strip()
        base = {"int32_t": "i32", "int": "i32", "int16_t": "i16", "int8_t": "i8", "int64_t": "i64", "float": "f32", "double": "f64",
                "uint8_t": "i8", "size_t": "size", "char": "char"}.get(base.split()[0] if base else "", "any")
        return ("ref", base)
    t = a.replace("const", "").split()
    base = t[0]
    return ("val", {"int32_t": "i32", "int": "i32", "size_t": "size", "float": "f32", "double": "f64", "int64_t": "i64"}.get(base, base))


_F_KIND = {"C_INT": "i32", "C_SHORT": "i16", "C_SIGNED_CHAR": "i8", "C_INT8_T": "i8", "C_INT64_T": "i64", "C_LONG_LONG": "i64",
           "C_FLOAT": "f32", "C_DOUBLE": "f64", "C_SIZE_T": "size", "C_INT32_T": "i32", "C_INT16_T": "i16"}


def _f_interfaces():
    lines = open(os.path.join(ROOT, "sosie_b200", "fortran", "sosie_gpu_c.f90")).read().split("\n")
    # join continuation lines, drop comments
    joined, cur = [], ""
    for ln in lines:
        ln = ln.split("!")[0].rstrip()
        if not ln.strip():
            continue
        s = ln.strip()
        if s.startswith("&"):
            s = s[1:].strip()
        if cur:
            cur += " " + s
        else:
            cur = s
        if cur.endswith("&"):
            cur = cur[:-1].rstrip()
            continue
        joined.append(cur); cur = ""
    out, i = {}, 0
    while i < len(joined):
        m = re.match(r"(?i)^(?:\w[\w\(\)]*\s+)?(FUNCTION|SUBROUTINE)\s+(\w+)\s*\(([^)]*)\)\s*BIND\s*\(\s*C\s*,\s*name\s*=\s*'(\w+)'\s*\)", joined[i])
        if not m:
            i += 1
            continue
        fname, args, cname = m.group(2), [a.strip().lower() for a in m.group(3).split(",") if a.strip()], m.group(4)
        decl = {}
        i += 1
        while i < len(joined) and not re.match(r"(?i)^END\s+(FUNCTION|SUBROUTINE)", joined[i]):
            d = joined[i]
            cm = re.match(r"(?i)^CHARACTER\s*\(\s*KIND\s*=\s*C_CHAR\s*\)\s*(.*?)::\s*(.*)$", d)
            if cm:
                for nm in _split_args(cm.group(2)):
                    decl[re.match(r"(\w+)", nm.strip()).group(1).lower()] = ("VALUE" in cm.group(1).upper(), "char")
                i += 1
                continue
            dm = re.match(r"(?i)^(TYPE\s*\(\s*(C_PTR|C_FUNPTR)\s*\)|INTEGER\s*\(\s*(\w+)\s*\)|REAL\s*\(\s*(\w+)\s*\))\s*(.*?)::\s*(.*)$", d)
            if dm:
                attrs, names = dm.group(5).upper(), dm.group(6)
                byval = "VALUE" in attrs
                if dm.group(2):
                    kind = "ptr"
                else:
                    kind = _F_KIND.get((dm.group(3) or dm.group(4)).upper(), "?")
                for nm in _split_args(names):
                    nm = nm.strip()
                    base = re.match(r"(\w+)", nm).group(1).lower()
                    decl[base] = (byval, kind)
            i += 1
        out[cname] = (fname, args, decl)
    return out


def test_every_fortran_interface_matches_a_header_prototype():
    protos, ifs = _c_prototypes(), _f_interfaces()
    assert len(ifs) >= 41, f"only {len(ifs)} interfaces parsed"
    problems = []
    for cname, (fname, args, decl) in ifs.items():
        if cname not in protos:
            problems.append(f"{cname}: not declared in include/sosie_b200.h")
            continue
        cargs = protos[cname]
        if len(cargs) != len(args):
            problems.append(f"{cname}: {len(args)} Fortran dummies, {len(cargs)} C parameters")
            continue
        for pos, (a, (how, ctype)) in enumerate(zip(args, cargs)):
            if a not in decl:
                problems.append(f"{cname}: dummy '{a}' has no declaration")
                continue
            byval, kind = decl[a]
            if how == "val":
                if not byval or kind != ctype:
                    problems.append(f"{cname} arg {pos + 1} '{a}': C takes {ctype} by value, Fortran passes {kind}{'' if byval else ' by reference'}")
            else:   # the C side takes a pointer
                if byval and kind != "ptr":
                    problems.append(f"{cname} arg {pos + 1} '{a}': C takes a pointer, Fortran passes {kind} by value")
                if not byval and kind != "ptr" and ctype not in ("any", kind):
                    problems.append(f"{cname} arg {pos + 1} '{a}': C element type {ctype}, Fortran {kind}")
    assert not problems, "\n".join(problems)


def _f_calls(path):
    """(name, number of actual arguments) of every sosie_* reference in a Fortran source (continuation lines joined)"""
    text = open(path).read().split("\n")
    joined, cur = [], ""
    for ln in text:
        ln = ln.split("!")[0].rstrip()
        s = ln.strip()
        if not s:
            continue
        if s.startswith("&"):
            s = s[1:].strip()
        cur = (cur + " " + s) if cur else s
        if cur.endswith("&"):
            cur = cur[:-1].rstrip()
            continue
        joined.append(cur); cur = ""
    calls = []
    for ln in joined:
        if re.match(r"(?i)^\s*(INTEGER|TYPE|REAL)\b.*\bFUNCTION\b", ln) or re.match(r"(?i)^\s*END\b", ln):
            continue
        for m in re.finditer(r"\b(sosie_\w+)\s*\(", ln):
            depth, j = 1, m.end()
            while j < len(ln) and depth:
                depth += ln[j] == "("
                depth -= ln[j] == ")"
                j += 1
            inner = ln[m.end():j - 1]
            calls.append((m.group(1), len([a for a in _split_args(inner) if a.strip()])))
    return calls


def test_drop_in_modules_call_the_interfaces_with_the_right_number_of_arguments():
    ifs = _f_interfaces()
    fdir = os.path.join(ROOT, "sosie_b200", "fortran")
    seen, problems = 0, []
    for fn in sorted(os.listdir(fdir)):
        if not fn.endswith(".f90"):
            continue
        for name, nargs in _f_calls(os.path.join(fdir, fn)):
            if name not in ifs:
                problems.append(f"{fn}: {name} has no interface in sosie_gpu_c.f90")
                continue
            seen += 1
            if nargs != len(ifs[name][1]):
                problems.append(f"{fn}: {name} called with {nargs} arguments, its interface has {len(ifs[name][1])}")
    assert seen >= 20, f"only {seen} calls found"
    assert not problems, "\n".join(problems)
