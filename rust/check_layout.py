#!/usr/bin/env python
"""cbindgen-style consistency check of the Rust shim against the C header, for a tree where rustc is absent.

  python rust/check_layout.py            exit status 0 = in step

Compares rust/src/image/writer/jpeg/cuda.rs with include/dmmt_cuda.h:
  * every `#[repr(C)] pub struct` -- field names, order and types -- against the C typedef of the same name
    (opaque handles: a zero-sized Rust struct against a forward-declared C struct);
  * every function of the `extern "C"` block -- return type, parameter count and parameter types -- against
    the C prototype of the same name;
  * every `pub const DMMT_*` against the `#define` (or dmmt_fmt enumerator) of the same name;
  * the applied patches only touch files that exist in the reference tree (when /root/reference is present).
The type map is the System V x86-64 / AArch64 LP64 one that both toolchains use.
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RS = os.path.join(ROOT, "rust", "src", "image", "writer", "jpeg", "cuda.rs")
HDR = os.path.join(ROOT, "include", "dmmt_cuda.h")

C_TO_RS = {
    "int": "c_int", "dmmt_fmt": "c_int", "size_t": "usize", "uint8_t": "u8", "uint16_t": "u16", "uint32_t": "u32",
    "uint64_t": "u64", "int16_t": "i16", "int32_t": "i32", "int64_t": "i64", "float": "f32", "char": "c_char",
    "void": "c_void",
}


def strip_comments(text, c_style):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    return text


def c_type_to_rust(decl):
    """'const dmmt_image *' -> '*const dmmt_image'; 'uint8_t **' -> '*mut *mut u8'; arrays decay to pointers."""
    decl = decl.strip()
    arr = decl.endswith("]")
    decl = re.sub(r"\[[^\]]*\]", "", decl).strip()
    stars = decl.count("*")
    words = [w for w in decl.replace("*", " ").split() if w]
    # `T *const *name`: a const AFTER a star qualifies that pointer, const before the base type qualifies the pointee
    base_const = "const" in words[: 2] and words[0] == "const"
    words = [w for w in words if w != "const"]
    base = words[0]
    if base == "struct":
        base = words[1]
    base = C_TO_RS.get(base, base)
    if arr:
        stars += 1
    out = base
    # innermost pointer carries the pointee's constness
    inner_const_pointers = len(re.findall(r"\*\s*const", decl))
    for i in range(stars):
        is_const = (i == 0 and base_const) or (i > 0 and i <= inner_const_pointers)
        out = ("*const " if is_const else "*mut ") + out
    return out


def parse_header():
    text = strip_comments(open(HDR).read(), True)
    consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(DMMT_\w+)\s+\(?(-?\d+)\)?", text)}
    m = re.search(r"typedef\s+enum\s*\{([^}]*)\}\s*dmmt_fmt", text)
    for part in m.group(1).split(","):
        k, v = part.split("=")
        consts[k.strip()] = int(v)
    structs = {}
    for m in re.finditer(r"typedef\s+struct\s*\{([^}]*)\}\s*(\w+)\s*;", text):
        fields = []
        for stmt in m.group(1).split(";"):
            stmt = stmt.strip()
            if not stmt:
                continue
            first, *rest = stmt.split(",")
            mm = re.match(r"(.*?)(\w+)\s*(\[[^\]]*\])?$", first.strip())
            ctype = mm.group(1)
            fields.append((mm.group(2), c_type_to_rust(ctype + (mm.group(3) or ""))))
            for more in rest:
                fields.append((more.strip().lstrip("*"), c_type_to_rust(ctype)))
        structs[m.group(2)] = fields
    opaque = set(re.findall(r"typedef\s+struct\s+(\w+)\s+\1\s*;", text))
    funcs = {}
    body = re.sub(r"typedef\s+(struct|enum)\s*\{[^}]*\}\s*\w+\s*;", " ", text)
    for m in re.finditer(r"([\w\s\*]+?)\b(dmmt_\w+)\s*\(([^)]*)\)\s*;", body):
        ret, name, params = m.group(1).strip().split("\n")[-1].strip(), m.group(2), m.group(3).strip()
        if "typedef" in ret or "#" in ret:
            continue
        plist = []
        if params and params != "void":
            for prm in params.split(","):
                prm = prm.strip()
                mm = re.match(r"(.*?)(\b\w+)?\s*(\[[^\]]*\])?$", prm)
                tname, pname, arr = mm.group(1), mm.group(2), mm.group(3) or ""
                if not tname.strip():       # unnamed parameter: the "name" is the type
                    tname, pname = pname, None
                elif pname and not re.search(r"[\*\s]$", tname) and pname not in ("const",):
                    # e.g. "dmmt_ctx *" handled by regexp; "const dmmt_options *" too
                    pass
                plist.append(c_type_to_rust((tname if pname is None else tname) + arr))
        r = "()" if ret == "void" else c_type_to_rust(ret)
        funcs[name] = (r, plist)
    return consts, structs, opaque, funcs


def parse_rust():
    text = strip_comments(open(RS).read(), False)
    consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"pub const (DMMT_\w+): c_int = (-?\d+);", text)}
    structs = {}
    for m in re.finditer(r"#\[repr\(C\)\]\s*pub struct (\w+)\s*\{([^}]*)\}", text):
        fields = []
        for f in m.group(2).split(","):
            f = f.strip()
            if not f:
                continue
            name, ty = f.split(":", 1)
            fields.append((name.replace("pub", "").strip(), ty.strip()))
        structs[m.group(1)] = fields
    funcs = {}
    ext = re.search(r'extern "C" \{(.*?)\n\}', text, flags=re.S).group(1)
    for m in re.finditer(r"pub fn (\w+)\s*\((.*?)\)\s*(->\s*([^;]+))?;", ext, flags=re.S):
        params = [p.split(":", 1)[1].strip() for p in m.group(2).split(",") if ":" in p]
        funcs[m.group(1)] = ((m.group(4) or "()").strip(), params)
    return consts, structs, funcs


def main():
    hc, hs, hopaque, hf = parse_header()
    rc, rs, rf = parse_rust()
    problems = []
    for name, val in rc.items():
        if name not in hc:
            problems.append(f"const {name}: not in the header")
        elif hc[name] != val:
            problems.append(f"const {name}: Rust {val}, header {hc[name]}")
    for name, fields in rs.items():
        if name in hopaque:
            if fields and fields != [("_opaque", "[u8; 0]")]:
                problems.append(f"struct {name}: opaque in the header, has fields in Rust")
            continue
        if name not in hs:
            problems.append(f"struct {name}: not in the header")
        elif hs[name] != fields:
            problems.append(f"struct {name}: header {hs[name]} != Rust {fields}")
    for name, (ret, params) in rf.items():
        if name not in hf:
            problems.append(f"fn {name}: not in the header")
            continue
        hret, hparams = hf[name]
        if hret != ret or hparams != params:
            problems.append(f"fn {name}: header {hret} {hparams} != Rust {ret} {params}")
    ref = "/root/reference"
    if os.path.isdir(ref):
        for pth in os.listdir(os.path.join(ROOT, "rust", "patches")):
            first = open(os.path.join(ROOT, "rust", "patches", pth)).readline().split()[1]
            if not os.path.exists(os.path.join(ref, first.split("/", 1)[1])):
                problems.append(f"patch {pth}: {first} is not a file of the reference tree")
    for p in problems:
        print("MISMATCH:", p)
    print(f"checked {len(rc)} constants, {len(rs)} structs, {len(rf)} functions of cuda.rs against include/dmmt_cuda.h: "
          f"{'OK' if not problems else str(len(problems)) + ' problem(s)'}")
    return 1 if problems else 0


if __name__ == "__main__":
    sys.exit(main())
