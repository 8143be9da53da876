"""Writes tests/golden/reference_interface.json: the public procedure signatures of the reference's Fortran modules
(src/array_utils.f90:11, src/lapack_wrapper.f90:9, src/davidson.f90:24, :273, :599), extracted with numpy.f2py's
Fortran parser (no compiler needed).  Run HERE, where /root/reference exists; the fixture travels, the reference does
not.  tests/test_abi_cpu.py::test_fortran_shim_interface_matches_the_reference parses fortran/davidson.f90 the same way
and compares.

  python tests/golden/make_reference_interface.py [/root/reference]
"""
import glob
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

MODULES = ("array_utils", "lapack_wrapper", "davidson_dense", "davidson_free", "davidson")


def crack(paths):
    from numpy.f2py import crackfortran
    crackfortran.verbose = 0
    crackfortran.quiet = 1
    # crackfortran keeps module-level state between calls
    crackfortran.reset_global_f2py_vars() if hasattr(crackfortran, "reset_global_f2py_vars") else None
    return crackfortran.crackfortran(list(paths))


def describe_var(v, nested):
    """One dummy argument (or function result) -> the part of its declaration a caller depends on."""
    if v is None or "typespec" not in v:
        return {"kind": "procedure" if nested else "undeclared"}
    d = {"type": v["typespec"]}
    ks = v.get("kindselector") or {}
    if ks:
        d["kind"] = str(ks.get("kind", ks.get("*", "")))
    cs = v.get("charselector") or {}
    if cs:
        d["len"] = str(cs.get("*", cs.get("len", "")))
    d["rank"] = len(v.get("dimension", []))
    if v.get("intent"):
        d["intent"] = sorted(x for x in v["intent"] if x in ("in", "out", "inout"))
    attrs = sorted(a for a in v.get("attrspec", []) if a in ("optional", "allocatable"))
    if attrs:
        d["attrs"] = attrs
    return d


def procedure_dummies(source, name):
    """dummy name -> abstract interface name, for `procedure(iface) :: a, b` declarations inside procedure `name`
    (f2py's parser skips that statement)."""
    m = re.search(r"(?:subroutine|function)\s+%s\s*\(.*?end\s+(?:subroutine|function)\s+%s\b" % (name, name), source, re.I | re.S)
    out = {}
    if m:
        for d in re.finditer(r"procedure\s*\(\s*(\w+)\s*\)\s*(?:,[^:]*)?::\s*([^\n!]*)", m.group(0), re.I):
            for a in d.group(2).split(","):
                out[a.strip().lower()] = d.group(1).lower()
    return out


def procedures(block, out, module=None, abstract=None, source=""):
    kind = block.get("block")
    if kind == "module":
        module = block["name"]
        abstract = {}
        for c in block.get("body", []):
            if c.get("block") == "abstract interface":
                for p in c.get("body", []):
                    abstract[p["name"]] = p
    if kind in ("subroutine", "function") and module in MODULES:
        nested = {}
        for a, iface in procedure_dummies(source, block["name"]).items():
            if a in block["args"] and abstract and iface in abstract:
                nested[a] = abstract[iface]
        for c in block.get("body", []):
            if c.get("block") in ("interface", "abstract interface"):
                for p in c.get("body", []):
                    if p.get("block") in ("subroutine", "function"):
                        nested[p["name"]] = p
        args = []
        for a in block["args"]:
            d = describe_var(block["vars"].get(a), a in nested)
            if a in nested:
                p = nested[a]
                d = {"kind": "procedure", "args": [dict(name=x, **describe_var(p["vars"].get(x), False)) for x in p["args"]]}
                res = p.get("result") or p["name"]
                if p["block"] == "function":
                    d["result"] = describe_var(p["vars"].get(res), False)
            args.append(dict(name=a, **d))
        entry = {"module": module, "block": kind, "args": args}
        if kind == "function":
            res = block.get("result") or block["name"]
            entry["result"] = describe_var(block["vars"].get(res), False)
        out.setdefault(block["name"], entry)
        return
    for c in block.get("body", []):
        procedures(c, out, module, abstract, source)


def public_names(path):
    """module -> names on its `public ::` statement (continuation lines joined)."""
    text = re.sub(r"&\s*\n\s*&?", " ", open(path).read())
    pub, module = {}, None
    for line in text.splitlines():
        m = re.match(r"\s*module\s+(\w+)\s*$", line, re.I)
        if m:
            module = m.group(1).lower()
        m = re.match(r"\s*public\s*::\s*(.*)$", line, re.I)
        if m and module:
            pub.setdefault(module, []).extend(x.strip().lower() for x in m.group(1).split(",") if x.strip())
    return pub


def generic_interfaces(path):
    """generic name -> specific procedures (e.g. generalized_eigensolver -> dense, free)."""
    out, cur = {}, None
    for line in open(path):
        m = re.match(r"\s*interface\s+(\w+)\s*$", line, re.I)
        if m:
            cur = m.group(1).lower()
            out[cur] = []
            continue
        if re.match(r"\s*end\s+interface", line, re.I):
            cur = None
        m = re.match(r"\s*(?:module\s+)?procedure\s+(?:::\s*)?(.*)$", line, re.I)
        if cur and m:
            out[cur].extend(x.strip().lower() for x in m.group(1).split(",") if x.strip())
    return out


def interface_of(paths):
    procs = {}
    source = "\n".join(open(p).read() for p in paths)
    for b in crack(paths):
        procedures(b, procs, source=source)
    pub, gen = {}, {}
    for p in paths:
        for k, v in public_names(p).items():
            pub.setdefault(k, []).extend(v)
        gen.update(generic_interfaces(p))
    return {"public": {k: sorted(set(v)) for k, v in pub.items() if k in MODULES}, "generic": gen, "procedures": procs}


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    paths = [os.path.join(ref, "src", f) for f in ("array_utils.f90", "lapack_wrapper.f90", "davidson.f90")]
    iface = interface_of(paths)
    keep = set()
    for names in iface["public"].values():
        keep |= set(names)
    for names in iface["generic"].values():
        keep |= set(names)
    iface["procedures"] = {k: v for k, v in sorted(iface["procedures"].items()) if k in keep}
    iface["source"] = "NLESC-JCER/Fortran_Davidson src/{array_utils,lapack_wrapper,davidson}.f90 via numpy.f2py.crackfortran"
    with open(os.path.join(HERE, "reference_interface.json"), "w") as fh:
        json.dump(iface, fh, indent=1, sort_keys=True)
        fh.write("\n")
    print("procedures:", ", ".join(iface["procedures"]))
    print("public:", iface["public"])
    print("generic:", iface["generic"])


if __name__ == "__main__":
    main()
