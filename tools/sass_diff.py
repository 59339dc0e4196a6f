"""Dev tool: which kernels of two builds of libdsgd.so differ, instruction for instruction (cuobjdump -sass; addresses and
encodings ignored).  Used to show that adding template variants leaves the kernels that were verified on the GPU untouched.
    python tools/sass_diff.py old/libdsgd.so distributed_sgd_b200/libdsgd.so
A kernel that gained trailing default template arguments (e.g. `..., 0>` -> `..., 0, 0>`) is matched by its name prefix."""
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], check=True, capture_output=True, text=True).stdout
    d, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            d[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?)\s*/\*", line)
        if m and cur is not None:
            d[cur].append(m.group(1))
    return d


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    differ = 0
    for name, body in sorted(a.items()):
        cands = [name] if name in b else [n for n in (name.replace("EEEv", "ELi0EEEv", 1), name.replace("EEEv", "ELi0ELi0EEEv", 1)) if n in b]
        if not cands:
            print("only in the first build:", name)
            differ += 1
        elif body != b[cands[0]]:
            print("DIFFERENT:", name, len(body), "->", len(b[cands[0]]), "instructions")
            differ += 1
    print(f"{len(a)} kernels in the first build, {len(b)} in the second, {differ} differ or are missing")
    return 1 if differ else 0


if __name__ == "__main__":
    raise SystemExit(main())
