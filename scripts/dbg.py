# scratch: phase timing of the CLI on cfg1 and hg38/16
import os, sys, subprocess, tempfile
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cli_bench as cb, bench
for name, chroms, cmd in [("cfg1", [("chr1", 10000000)], ["--chromosomes=g.chroms", "--novalue", "=", "sum", "--window=101", "=", "localmax", "--neighborhood=11"]),
                          ("div16", bench.scaled_genome(16), ["--chromosomes=g.chroms", "--novalue", "--precision=3", "=", "smooth", "--window=101"])]:
    with tempfile.TemporaryDirectory() as d:
        reads = cb.write_case(d, chroms, 1)
        env = dict(os.environ, GENODSP_TIMING="1")
        for rep in range(2):
            with open(reads, "rb") as fin:
                p = subprocess.run([cb.OURS] + cmd, stdin=fin, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, cwd=d, env=env)
            print(name, rep, p.stderr.decode().replace("\n", " | "), flush=True)
