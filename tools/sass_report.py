"""Blackwell-specific SASS in the shipped library: `python tools/sass_report.py > profiles/sass_r02.txt`.
cuobjdump -sass of worldrenderer_b200/lib/libwr_b200.so, one line per (mnemonic, kernel) with its count."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "worldrenderer_b200", "lib", "libwr_b200.so")
WHAT = [
    (r"UTMALDG[\w.]*", "TMA bulk tensor load (cp.async.bulk.tensor)"),
    (r"SYNCS[\w.]*", "mbarrier arrive / expect-tx / try_wait (TMA completion)"),
    (r"LDGMC[\w.]*", "multimem.ld_reduce (in-switch sum over the NVSwitch multicast window)"),
    (r"ACQBULK", "griddepcontrol.wait (programmatic dependent launch)"),
    (r"REDG\.E\.MIN\.64[\w.]*", "64-bit atomicMin without return: visibility resolve (depth key << 32 | id)"),
    (r"REDG\.E\.ADD\.64[\w.]*", "64-bit integer atomic add: exact fixed-point vertex-normal splat"),
    (r"VIMNMX3?\.U16x2[\w.]*", "packed 16-bit min / max of three: bounding box of the compact vertex records"),
    (r"VIMNMX3(\.U32|\.S32)?\b", "three-input integer min / max"),
]
sass = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
cur, hits = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    for pat, _ in WHAT:
        for h in re.findall(r"\b(" + pat + r")", line):
            hits[h if isinstance(h, str) else h[0]][cur] += 1
def short(name):
    out = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    out = re.sub(r"\(anonymous namespace\)::", "", out)
    return re.sub(r"\(.*", "", out).replace("void ", "")
print(f"# cuobjdump -sass worldrenderer_b200/lib/libwr_b200.so   (architectures in the fatbin: {', '.join(archs)})")
print("# nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -lineinfo")
for pat, what in WHAT:
    names = [h for h in hits if re.fullmatch(pat, h)]
    if not names:
        print(f"\n## {pat}: none")
        continue
    print(f"\n## {what}")
    for h in sorted(names):
        for k, n in sorted(hits[h].items(), key=lambda kv: -kv[1]):
            print(f"{h:36s} x{n:<3d} {short(k)}")
