"""Regenerates the `extern "C"` block of integration/sarpro-gpu-sys/src/lib.rs from include/sarpro_gpu.h (struct
definitions are hand-written above the block). Usage: python integration/gen_sys.py [--check]"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TMAP = {"int": "c_int", "size_t": "usize", "float": "f32", "double": "f64", "uint8_t": "u8", "uint16_t": "u16", "uint32_t": "u32",
        "uint64_t": "u64", "void": "c_void", "char": "c_char", "sarpro_ctx": "sarpro_ctx", "sarpro_stats": "sarpro_stats",
        "sarpro_band": "sarpro_band", "sarpro_image": "sarpro_image", "sarpro_resize_meta": "sarpro_resize_meta",
        "sarpro_timing": "sarpro_timing", "sarpro_scene": "sarpro_scene", "sarpro_batch_report": "sarpro_batch_report"}


def rtype(t):
    t = t.strip()
    const = t.startswith("const ")
    if const:
        t = t[6:].strip()
    stars = t.count("*")
    r = TMAP[t.replace("*", "").strip()]
    for i in range(stars):
        r = ("*const " if (const and i == 0) else "*mut ") + r
    return r


def prototypes():
    h = open(os.path.join(ROOT, "include", "sarpro_gpu.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return [" ".join(p.split()) for p in re.findall(r"\n((?:int|void|const char\*|void\*|size_t)\s+\*?sarpro_[a-z0-9_]+\s*\([^;]*?\));", h)]


def rust_decls():
    out = []
    for p in prototypes():
        m = re.match(r"(.+?)\s*\*?\s*(sarpro_[a-z0-9_]+)\s*\((.*)\)$", p)
        ret, name, args = m.group(1), m.group(2), m.group(3)
        if p.startswith("const char*"):
            ret = "const char*"
        if p.startswith("void*"):
            ret = "void*"
        al = []
        if args.strip() != "void":
            for a in args.split(","):
                mm = re.match(r"(.+?)([a-z_0-9]+)$", a.strip())
                ty, nm = mm.group(1).strip(), mm.group(2)
                if nm in ("type", "in", "ref", "or"):
                    nm += "_"
                al.append(f"{nm}: {rtype(ty)}")
        out.append(f"    pub fn {name}({', '.join(al)}){'' if ret == 'void' else ' -> ' + rtype(ret)};")
    return out


def main():
    path = os.path.join(ROOT, "integration", "sarpro-gpu-sys", "src", "lib.rs")
    src = open(path).read()
    head, _, _ = src.partition('extern "C" {\n')
    new = head + 'extern "C" {\n' + "\n".join(rust_decls()) + "\n}\n"
    if "--check" in sys.argv:
        sys.exit(0 if new == src else 1)
    open(path, "w").write(new)


if __name__ == "__main__":
    main()
