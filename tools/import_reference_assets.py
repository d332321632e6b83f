#!/usr/bin/env python3
"""Re-emits the reference's scene DATA (OBJ/MTL meshes under resources/models) in canonical form into
resources/models/.  These are input fixtures of the parity tests and of BASELINE.json's Cornell-box
configs, not code: comments and unused statements (vt, s, Ns/Ka/Ks/illum) are dropped, numbers are
kept verbatim so the parsed floats are bit-identical.  Run in the build container (needs
/root/reference); the outputs are committed."""
import os
import sys

REF = "/root/reference/resources/models"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "resources", "models")


def convert_obj(src, dst):
    out = ["# mesh data from lens_trace resources/models/%s (canonical re-emit, see tools/import_reference_assets.py)"
           % os.path.basename(src)]
    for line in open(src):
        t = line.split()
        if not t or t[0].startswith("#"):
            continue
        if t[0] in ("mtllib", "o", "v", "vn", "usemtl"):
            out.append(" ".join(t))
        elif t[0] == "f":
            corners = []
            for c in t[1:]:
                parts = c.split("/")
                v = parts[0]
                vn = parts[2] if len(parts) > 2 else ""
                corners.append("%s//%s" % (v, vn) if vn else v)
            out.append("f " + " ".join(corners))
    open(dst, "w").write("\n".join(out) + "\n")


def convert_mtl(src, dst):
    out = ["# materials from lens_trace resources/models/%s (canonical re-emit)" % os.path.basename(src)]
    for line in open(src):
        t = line.split()
        if not t or t[0].startswith("#"):
            continue
        if t[0] == "newmtl":
            out.append("")
            out.append(" ".join(t))
        elif t[0] in ("Kd", "Ke", "Ni", "d", "Tr"):
            out.append(" ".join(t))
    open(dst, "w").write("\n".join(out) + "\n")


def main():
    if not os.path.isdir(REF):
        sys.exit("reference not present at %s" % REF)
    os.makedirs(OUT, exist_ok=True)
    for name in sorted(os.listdir(REF)):
        if name.endswith(".obj"):
            convert_obj(os.path.join(REF, name), os.path.join(OUT, name))
        elif name.endswith(".mtl"):
            convert_mtl(os.path.join(REF, name), os.path.join(OUT, name))
        print("wrote", name)


if __name__ == "__main__":
    main()
