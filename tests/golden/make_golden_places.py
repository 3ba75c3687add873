"""Adds tests/golden/<case>/reference_step3_places.txt: what the UNMODIFIED reference binary (oracle/_ref/w2rap-contigger) prints about
RepathInMemory's `places` (paths/long/large/Repath.cc:46-72) when step 3 is run on the committed step-2 outputs of each golden case,
for several large K: "<K2> <paths> <places kept> <unique places>".  The reference does not write the places themselves anywhere; the
two counts per K2 are what pins the oracle's restatement (tests/test_oracle.py::test_places_counts_match_the_reference_log).
Run here (CPU container); /root/reference does not exist on the GPU box."""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import w2r_testlib as T  # noqa: E402


def main():
    for case in ("circ", "rich", "long"):
        src = os.path.join(HERE, case)
        lines = []
        for K2 in (80, 100, 132, 160, 200, 260, 320):
            d = tempfile.mkdtemp(prefix="w2rap_places_")
            for f in os.listdir(src):
                shutil.copy(os.path.join(src, f), d)
            r = subprocess.run([T.REF_BIN, "-t", "2", "-o", d, "-p", "x", "-r", "a.fq,b.fq", "--from_step", "3", "--to_step", "3", "-K", str(K2)],
                               capture_output=True, text=True)
            out = r.stdout
            m1 = re.search(r"constructing places from (\d+) paths", out)
            m2 = re.search(r"sorting (\d+) places", out)
            m3 = re.search(r"(\d+) unique places", out)
            shutil.rmtree(d, ignore_errors=True)
            if not (m1 and m2 and m3):
                print(case, K2, "no places lines (rc %d)" % r.returncode, out[-300:], file=sys.stderr)
                continue
            lines.append("%d %s %s %s" % (K2, m1.group(1), m2.group(1), m3.group(1)))
        with open(os.path.join(src, "reference_step3_places.txt"), "w") as f:
            f.write("\n".join(lines) + "\n")
        print(case, lines)


if __name__ == "__main__":
    main()
