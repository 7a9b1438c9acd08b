#!/usr/bin/env python
"""Write profiles/ncu_traffic.json from an ncu --set full capture of the label kernel, stamped with the digest of
the kernel sources it was captured from (bench.py only quotes it when the digest still matches).
usage: ncu_traffic.py report.ncu-rep kmers_per_launch "how it was captured" """
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rep, nk, how = sys.argv[1], float(sys.argv[2]), sys.argv[3]
    import bench

    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    d = {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}

    def b(k):
        return float(d[k][0].replace(",", "")) * scale[d[k][1]]

    rd, wr = b("dram__bytes_read.sum"), b("dram__bytes_write.sum")
    j = {"kernel": d["Kernel Name"][0], "dram_bytes_per_kmer": (rd + wr) / nk, "dram_bytes_read": rd, "dram_bytes_write": wr,
         "kmers_per_launch": int(nk), "source_digest": bench.source_digest(), "source": how}
    json.dump(j, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(j))


if __name__ == "__main__":
    main()
