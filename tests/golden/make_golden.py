"""Regenerates tests/golden/* by running the UNMODIFIED reference binary (oracle/_ref/w2rap-contigger, built from
/root/reference by `make -C oracle`) on small seeded synthetic read sets.  Run here (CPU container); the outputs are
committed because /root/reference does not exist on the GPU box.

  step1/   : FASTQ pair -> reference step 1 -> frag_reads_orig.fastb/.qualp (pins the feudal + PQVec formats against
             files written by the reference's own encoder) + the quals/bases the FASTQ held (expected.npz)
  circ/    : 12 circular replicons (odd/even lengths) + planted palindromes -> reference step 2 (-t 1)
  rich/    : repeats + SNP haplotype + palindromes + plasmid, variable read lengths -> reference step 2 (-t 1)
  long/    : the same kind of genome with reads of up to 600 bases -> reference step 2 (-t 1)
Each step-2 case holds the input read stores and the reference's x.small_K.hbv / x.small_K.paths / small_K.freqs.
"""
import os
import shutil
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import w2r_testlib as T  # noqa: E402


def keep_only(d, names):
    for f in os.listdir(d):
        if f not in names:
            p = os.path.join(d, f)
            shutil.rmtree(p) if os.path.isdir(p) else os.remove(p)


def step2_case(name, rs):
    d = os.path.join(HERE, name)
    shutil.rmtree(d, ignore_errors=True)
    T.write_fastb_qualp(d, rs)
    out, _ = T.run_reference_step2(d, threads=1)
    with open(os.path.join(d, "reference_stdout.txt"), "w") as f:
        f.write("\n".join(l.split(": ", 1)[-1] for l in out.splitlines() if "kmers" in l or "edges" in l or "pathed" in l) + "\n")
    keep_only(d, {"frag_reads_orig.fastb", "frag_reads_orig.qualp", "x.small_K.hbv", "x.small_K.paths", "small_K.freqs",
                  "reference_stdout.txt"})


def main():
    # ---- step1
    d = os.path.join(HERE, "step1")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    rng = np.random.default_rng(1)
    G = rng.integers(0, 4, 5000)
    seqs, quals = [], []
    for i in range(150):
        p = int(rng.integers(0, 4400))
        frag = G[p:p + 500]
        for k in range(2):
            L = int(rng.integers(40, 251))
            s = (frag if k == 0 else 3 - frag[::-1])[:L]
            mode = i % 3
            if mode == 0:
                q = np.concatenate([np.full(L - 20, 37), rng.integers(2, 41, 20)])
            elif mode == 1:
                q = rng.integers(0, 42, L)
            else:
                q = np.clip(37 - (np.arange(L) // 40) * 5 + rng.integers(-1, 2, L), 2, 41)
            seqs.append(s.astype(np.uint8)); quals.append(q.astype(np.uint8))
    for k in range(2):
        with open(os.path.join(d, "r%d.fastq" % (k + 1)), "w") as f:
            for i in range(k, len(seqs), 2):
                f.write("@r%d\n%s\n+\n%s\n" % (i // 2, "".join("ACGT"[b] for b in seqs[i]), "".join(chr(33 + x) for x in quals[i])))
    subprocess.run([T.REF_BIN, "-t", "1", "-o", d, "-p", "x", "-r", os.path.join(d, "r1.fastq") + "," + os.path.join(d, "r2.fastq"),
                    "--to_step", "1"], check=True, stdout=subprocess.DEVNULL)
    np.savez_compressed(os.path.join(d, "expected.npz"), lens=np.array([len(s) for s in seqs], np.uint32),
                        bases=np.concatenate(seqs), quals=np.concatenate(quals))
    keep_only(d, {"frag_reads_orig.fastb", "frag_reads_orig.qualp", "expected.npz"})

    # ---- circ
    rng = np.random.default_rng(11)
    reps = [(rng.integers(0, 4, int(L), dtype=np.uint8), True, float(L)) for L in [300, 301, 402, 517, 1000, 1001, 2048, 777, 64, 61, 130, 131]]
    reps.append((T.make_genome(rng, 5000, 0, 3), False, 5000.0))
    tot = sum(r[2] for r in reps)
    step2_case("circ", T.flatten_reads(*T.simulate_reads(rng, reps, int(tot * 80 // 500), 250, frag_mean=300, frag_sd=20), pq_mode=1))

    # ---- long: 600-base reads of varying length (several 192-k-mer map tiles per read, many gaps per read)
    step2_case("long", T.rich_set(seed=9, genome=20000, cov=30, read_len=600, families=3, palindromes=2, plasmid=1500, vary_len=True, pq_mode=1))

    # ---- rich
    step2_case("rich", T.rich_set(seed=5, genome=40000, cov=50, families=5, palindromes=4, plasmid=1500, vary_len=True, pq_mode=1))


if __name__ == "__main__":
    main()
