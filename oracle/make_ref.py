"""Recipe for oracle/_ref/ (test / baseline infrastructure, NOT product code): the UNMODIFIED reference sources of the
hot path, taken from where they lie under /root/reference, plus the 2-file pygame stub the reference needs to import
without a display.

    python oracle/make_ref.py            # build container only (needs /root/reference); idempotent

The reference is pure Python - there is nothing to compile: "building" it means placing its own files next to a stub
for the one import that is absent here (pygame, imported unconditionally at env/env_small.py:13-14 and only used when
gamemode == 'pygame').  oracle/_ref/ is git-ignored (no reference source ever enters the history) but NOT gpurun-
ignored, so it travels to the GPU box, where /root/reference does not exist; bench.py's `--impl reference` arm and its
`cpu_baseline` leg import the reference from there (kind "reference"), falling back to the oracle port
(oracle/omok_oracle.py, kind "port") only when oracle/_ref/ is missing.  MANIFEST.json records the sha256 of every file
so a run can prove that what it timed is byte-identical to the upstream source.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("ALPHA_OMOK_REF", "/root/reference/2_AlphaOmok")
OUT = os.path.join(HERE, "_ref")
FILES = ["agents.py", "utils.py", "model.py", "env/env_small.py", "env/env_regular.py"]  # the path of SURVEY 8(a)


def sha256(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(verbose=True):
    """Returns OUT if the reference is present (and now assembled), None if /root/reference is absent (GPU box)."""
    if not os.path.isdir(REF_SRC):
        return OUT if os.path.exists(os.path.join(OUT, "MANIFEST.json")) else None
    dst = os.path.join(OUT, "2_AlphaOmok")
    os.makedirs(os.path.join(dst, "env"), exist_ok=True)
    manifest = {"source": REF_SRC, "files": {}}
    for rel in FILES:
        shutil.copyfile(os.path.join(REF_SRC, rel), os.path.join(dst, rel))
        manifest["files"][rel] = sha256(os.path.join(dst, rel))
    stub = os.path.join(OUT, "stubs", "pygame")
    os.makedirs(stub, exist_ok=True)
    with open(os.path.join(stub, "__init__.py"), "w") as f:
        f.write("# stub: the reference imports pygame unconditionally (env/env_small.py:13) but only draws with it in\n"
                "# gamemode == 'pygame'; text mode never touches it\n")
    with open(os.path.join(stub, "locals.py"), "w") as f:
        f.write("QUIT = 12\n")
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    if verbose:
        print("oracle/_ref assembled from", REF_SRC, "(%d files)" % len(FILES))
    return OUT


def verify():
    """True iff every file under oracle/_ref/2_AlphaOmok still has the recorded hash."""
    try:
        with open(os.path.join(OUT, "MANIFEST.json")) as f:
            manifest = json.load(f)
        return all(sha256(os.path.join(OUT, "2_AlphaOmok", rel)) == h for rel, h in manifest["files"].items())
    except Exception:
        return False


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
