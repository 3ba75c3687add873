# GPU check of host/step12_main.cc: FASTQ pair -> step files; the step-2 files must equal what the (tested) `step2` program writes
# from the step-1 files step12 left behind.
set -e
D=$(mktemp -d)
python - "$D" <<'P'
import os, sys
sys.path.insert(0, "tests")
import numpy as np
import w2r_testlib as T
rng = np.random.default_rng(4)
g = T.make_genome(rng, 30000, 2, 1)
codes, quals, lens = T.simulate_reads(rng, [(g, False, 1.0)], 30000 * 40 // 500, 250)
d = sys.argv[1]
for k in range(2):
    with open(os.path.join(d, "r%d.fastq" % (k + 1)), "w") as f:
        for i in range(k, len(lens), 2):
            L = int(lens[i])
            f.write("@r%d\n%s\n+\n%s\n" % (i // 2, "".join("ACGT"[b] for b in codes[i][:L]), "".join(chr(33 + int(q)) for q in quals[i][:L])))
P
mkdir -p $D/out
./w2rap-contigger_b200/step12 $D/r1.fastq $D/r2.fastq $D/out x | tail -3
./w2rap-contigger_b200/step2 $D/out y --quiet | tail -1
cmp $D/out/x.small_K.hbv $D/out/y.small_K.hbv && cmp $D/out/x.small_K.paths $D/out/y.small_K.paths && echo "step12 == step2 on the same stores: OK ($(stat -c %s $D/out/x.small_K.hbv) B hbv, $(stat -c %s $D/out/x.small_K.paths) B paths)"
rm -rf $D
