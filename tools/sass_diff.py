#!/usr/bin/env python
"""Compares the device code of the current libmeso_b200.so with the one built from another commit, kernel by kernel
(instruction streams from `cuobjdump -sass`, addresses and encodings stripped):
    python tools/sass_diff.py <commit>
    python tools/sass_diff.py --digest [file]     # print (or check against file) one md5 per kernel of the current library
Used when host-side or flag-guarded changes are made without a GPU at hand: every kernel whose SASS is unchanged behaves as
it did when that commit was validated on hardware."""
import hashlib
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernels(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    d, cur = {}, None
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            d[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            d[cur].append(re.sub(r"/\*[0-9a-f]+\*/", "", line).strip())
    return {k: hashlib.md5("\n".join(v).encode()).hexdigest() for k, v in d.items()}


def digest(path=None):
    """kernel -> md5 of its instruction stream for the library in the tree; with a file: compare, report what changed"""
    b = kernels(os.path.join(ROOT, "meso_b200", "libmeso_b200.so"))
    if path is None:
        for k in sorted(b):
            print(b[k], k)
        return 0
    a = dict((line.split()[1], line.split()[0]) for line in open(path) if len(line.split()) == 2)
    changed = sorted(k for k in a if k in b and a[k] != b[k])
    gone, new = sorted(k for k in a if k not in b), sorted(k for k in b if k not in a)
    print("%s: %d kernels, library: %d; changed: %d, gone: %d, new: %d" % (path, len(a), len(b), len(changed), len(gone), len(new)))
    for k in changed + gone + new:
        print("  " + ("changed" if k in changed else "gone" if k in gone else "new"), k)
    return 1 if changed or gone or new else 0


def main():
    if sys.argv[1] == "--digest":
        return digest(sys.argv[2] if len(sys.argv) > 2 else None)
    commit = sys.argv[1]
    with tempfile.TemporaryDirectory() as tmp:
        wt = os.path.join(tmp, "wt")
        subprocess.check_call(["git", "-C", ROOT, "worktree", "add", "-q", wt, commit])
        try:
            subprocess.check_call(["make", "-C", os.path.join(wt, "meso_b200", "csrc"), "-j8", "-s"])
            a = kernels(os.path.join(wt, "meso_b200", "libmeso_b200.so"))
        finally:
            subprocess.call(["git", "-C", ROOT, "worktree", "remove", "--force", wt])
    b = kernels(os.path.join(ROOT, "meso_b200", "libmeso_b200.so"))
    changed = sorted(k for k in a if k in b and a[k] != b[k])
    print("%s: %d kernels, now: %d; changed: %d, gone: %d, new: %d" % (commit, len(a), len(b), len(changed), len([k for k in a if k not in b]),
                                                                       len([k for k in b if k not in a])))
    for k in changed:
        print("  changed", k)
    gone, new = sorted(k for k in a if k not in b), sorted(k for k in b if k not in a)
    for k in gone:
        same = [n for n in new if b[n] == a[k]]
        print("  gone   ", k, ("-> identical code now named " + same[0]) if same else "")
    for k in new:
        if not any(a[g] == b[k] for g in gone):
            print("  new    ", k)
    return 1 if changed else 0


if __name__ == "__main__":
    sys.exit(main())
