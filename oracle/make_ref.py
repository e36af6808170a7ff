"""Recipe that makes the UNMODIFIED reference phase-space path runnable on the GPU box -- TEST INFRASTRUCTURE ONLY.

The reference's path is pure Python (memflow/phasespace/{phasespace,rambo_generator,utils}.py; they need torch + numpy
only), so "building" it means staging those three files, byte for byte, under the git-ignored
`oracle/_ref/` (never committed; like the built .so files it travels to the GPU box with the gpurun snapshot, where
/root/reference does not exist). `phasespace.py:4` imports the `particle` package without using it: a two-line stub
package is written next to the copies (the same stub tests/golden/make_golden.py injects).

    python oracle/make_ref.py            # in the build container; __graft_entry__.build() runs it when /root/reference exists

bench.py's `cpu_baseline.reference_torch` leg and `--impl reference` import the staged package through
oracle/ref_torch.py; nothing under memflow_b200/ ever does.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MEMFLOW_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ("memflow/phasespace/phasespace.py", "memflow/phasespace/rambo_generator.py", "memflow/phasespace/utils.py")


def make_ref(quiet=False):
    """Returns True when oracle/_ref is in place (staged now or earlier), False when there is no reference tree here."""
    if not os.path.isdir(os.path.join(REF, "memflow", "phasespace")):
        return os.path.exists(os.path.join(DST, "MANIFEST.json"))
    manifest = {"source": REF, "files": {}}
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
        manifest["files"][rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    for pkg in ("memflow", "memflow/phasespace"):   # empty package markers (the reference's own are empty too)
        open(os.path.join(DST, pkg, "__init__.py"), "w").close()
    os.makedirs(os.path.join(DST, "particle"), exist_ok=True)
    with open(os.path.join(DST, "particle", "__init__.py"), "w") as f:
        f.write('"""Stub of the `particle` package: memflow/phasespace/phasespace.py imports it and never uses it."""\n'
                "class Particle:\n    pass\n")
    json.dump(manifest, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1)
    if not quiet:
        print("oracle/_ref staged from %s: %s" % (REF, ", ".join(FILES)))
    return True


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)
