"""SASS evidence for the judge (B200_PROFILING.md, "What proves a Blackwell-native kernel"): per kernel of libsfmb200.so the counts
of the mnemonics that matter -- UTC*MMA (tcgen05.mma), LDTM (tcgen05.ld), UBLKCP (cp.async.bulk), UTCBAR / SYNCS (mbarrier traffic),
DMMA (fp64 mma.sync), VIMNMX3 (3-input integer min/max), LDGSTS (cp.async), RED / ATOM (atomics), plus registers from the
ptxas logs.  Runs without a GPU:   python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "sfm-toy-library_b200", "lib", "libsfmb200.so")
WATCH = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTCBAR", "SYNCS", "DMMA", "HMMA", "VIMNMX3", "VIMNMX", "LDGSTS",
         "REDG", "ATOM", "ATOMS", "ATOMG", "MATCH", "VOTE", "POPC", "MUFU", "DFMA", "DADD", "DMUL"]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except Exception:
        return name


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); kernels[cur] = collections.Counter(); continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(2).split(".")[0]
            kernels[cur]["_total"] += 1
            if op in WATCH:
                kernels[cur][op] += 1
                if op == "REDG" and ".F64" in line:
                    kernels[cur]["REDG.F64"] += 1
    print("# cuobjdump -sass sfm-toy-library_b200/lib/libsfmb200.so  (sm_100a); counts of static instructions per kernel")
    print("# tcgen05.mma -> UTCIMMA (kind::i8), tcgen05.ld -> LDTM, cp.async.bulk -> UBLKCP, mma.sync f64 -> DMMA, cp.async -> LDGSTS")
    for k, c in kernels.items():
        name = demangle(k)
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name)
        items = ", ".join(f"{op} {c[op]}" for op in WATCH + ["REDG.F64"] if c[op])
        print(f"{name:70s} instr {c['_total']:6d}  {items}")


if __name__ == "__main__":
    sys.exit(main())
