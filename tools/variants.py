"""Builds libwr_b200 variants with different -D switches and times config B with each (stage table, us).

    python tools/variants.py build  name1:-DWR_X=1,-DWR_Y=2  name2:...     (here, no GPU needed)
    python tools/variants.py run [--mesh terrain|sphere] [--depth controlnet]   (on the GPU box)

Variant libraries go to worldrenderer_b200/lib/variants/ (git-ignored, shipped by gpurun)."""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "worldrenderer_b200", "lib", "variants")


def build(specs):
    from worldrenderer_b200 import build_native as bn
    os.makedirs(VDIR, exist_ok=True)
    for old in glob.glob(os.path.join(VDIR, "*.so")):
        os.remove(old)
    procs = []
    for spec in specs:
        name, _, flags = spec.partition(":")
        out = os.path.join(VDIR, f"lib_{name}.so")
        cmd = [bn._nvcc(), *bn.NVCC_FLAGS, *[f for f in flags.split(",") if f], *bn.sources(), "-o", out]
        procs.append((name, subprocess.Popen(cmd)))
    for name, p in procs:
        if p.wait() != 0:
            raise SystemExit(f"variant {name} failed to build")
        print("built", name)


def run(extra):
    libs = sorted(glob.glob(os.path.join(VDIR, "*.so")))
    for lib in libs:
        env = dict(os.environ, WR_B200_LIB=lib)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "prof_render.py"), "--per-view", "--quick", "--steps", "5", *extra],
                           env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("all views")]
        print(os.path.basename(lib), line[0] if line else ("FAILED " + r.stderr[-400:]))


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        run(sys.argv[2:])
