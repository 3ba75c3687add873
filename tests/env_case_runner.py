"""Runs one product-vs-oracle comparison in a fresh process (the library reads its diagnostic environment switches once per
process, so cases that need them cannot share the pytest process).  usage: env_case_runner.py genome cov seed"""
import sys

import w2r_testlib as T


def main():
    genome, cov, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    rs = T.rich_set(seed=seed, genome=genome, cov=cov, families=4, palindromes=2, plasmid=1500)
    want = T.run_oracle(rs, T.default_params(dump_kmers=2, apply_fixpaths=1))
    got = T.run_product(rs, T.default_params(dump_kmers=2, apply_fixpaths=1))
    T.assert_graph_equal(want, got)
    print("OK launches=%d passes=%d" % (got["timings"]["kernel_launches"], got["timings"]["count_passes"]))


if __name__ == "__main__":
    main()
