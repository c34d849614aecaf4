"""Per-kernel SASS comparison of two builds of libnessai_b200.so (no GPU needed):
    python scripts/sass_diff.py OLD.so [NEW.so]
prints which kernels are byte-identical (instruction addresses stripped), changed, added, removed.
Used to show that a refactor (moving a kernel into a header, guarding inline PTX for the CPU SIMT
shim of tests/_hostcheck) leaves the code that ran on the GPU untouched."""

import hashlib
import os
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    res, cur, buf = {}, None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if cur:
                res[cur] = hashlib.md5("\n".join(buf).encode()).hexdigest()
            cur, buf = m.group(1), []
        elif cur:
            buf.append(re.sub(r"/\*[0-9a-f]{4}\*/", "", line))
    if cur:
        res[cur] = hashlib.md5("\n".join(buf).encode()).hexdigest()
    return res


def main():
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    old = kernels(sys.argv[1])
    new = kernels(sys.argv[2] if len(sys.argv) > 2 else os.path.join(here, "nessai_b200", "lib", "libnessai_b200.so"))
    print(f"kernels: {len(old)} -> {len(new)}")
    print("byte-identical:", sum(1 for k in old if new.get(k) == old[k]))
    for title, names in (("changed", [k for k in old if k in new and new[k] != old[k]]),
                         ("removed / re-signed", [k for k in old if k not in new]),
                         ("added", [k for k in new if k not in old])):
        print(f"{title}:")
        for k in names:
            print("   ", subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()[:110])


if __name__ == "__main__":
    main()
