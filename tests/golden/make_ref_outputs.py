"""Golden OUTPUTS of the reference's own Python, run here (in the container that has /root/reference) on the test
shims of its third-party natives (tests/refshim/, CPU oracle backend):

  dory_subset_abund.json   count_dominator_abundance.main (search/count_dominator_abundance.py:24-170) over
                           dory-subset.fa: per-unitig abundances + node sizes (cdbg_abund.csv) and the per-dominator
                           table (dom_abund.csv), as the reference wrote them

The GPU tests compare the CUDA path with these vectors; tests/test_refshim_cpu.py re-runs the reference and checks
that the committed file is what it produces.  usage: python tests/golden/make_ref_outputs.py"""
import csv
import gzip
import json
import os
import shutil
import sys
import tempfile
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def reference_modules():
    """the reference package without its version check, on the shims (oracle backend)"""
    os.environ["SGC_REFSHIM_BACKEND"] = "oracle"
    for p in (os.path.join(ROOT, "tests", "refshim"), ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    pkg = types.ModuleType("spacegraphcats")
    pkg.__path__ = [os.path.join(REF, "spacegraphcats")]
    sys.modules["spacegraphcats"] = pkg
    import spacegraphcats.cdbg.hashing as hashing
    import spacegraphcats.cdbg.index_cdbg_by_kmer as index_cdbg_by_kmer
    import spacegraphcats.search.count_dominator_abundance as cda

    return hashing, index_cdbg_by_kmer, cda


def run_abundance(workdir):
    """-> dict(cdbg=[[id, abund, size], ...], dom=[[id, level, abund, size], ...]) from the reference's CSV files"""
    hashing, index_cdbg_by_kmer, cda = reference_modules()  # (also puts the repo root on sys.path)
    from oracle import oracle as O

    recs = O.read_unitigs_tsv(os.path.join(HERE, "dory_k21_unitigs.tsv.gz"))
    cdbg, catlas = os.path.join(workdir, "dory_k21"), os.path.join(workdir, "dory_k21_r1")
    os.makedirs(cdbg)
    os.makedirs(catlas)
    table, sizes = index_cdbg_by_kmer.build_mphf(21, lambda: iter(recs))
    hashing.MPHF_KmerIndex(21, table, sizes).save(cdbg)
    shutil.copy(os.path.join(HERE, "dory_k21_contigs.info.csv"), os.path.join(cdbg, "contigs.info.csv"))
    for f in ("catlas.csv", "first_doms.txt"):
        shutil.copy(os.path.join(HERE, "dory_k21_r1", f), os.path.join(catlas, f))
    sample = os.path.join(workdir, "dory-subset.fa")
    with gzip.open(os.path.join(HERE, "data", "dory-subset.fa.gz"), "rb") as fi, open(sample, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    outdir = os.path.join(workdir, "abund")
    assert cda.main([cdbg, catlas, sample, "--outdir", outdir]) == 0
    with open(os.path.join(outdir, "dory-subset.fa.cdbg_abund.csv"), newline="") as fp:
        cdbg_rows = [[int(x) for x in r] for r in list(csv.reader(fp))[1:]]
    with open(os.path.join(outdir, "dory-subset.fa.dom_abund.csv"), newline="") as fp:
        dom_rows = [[int(x) for x in r] for r in list(csv.reader(fp))[1:]]
    return {"cdbg": cdbg_rows, "dom": dom_rows}


def main():
    with tempfile.TemporaryDirectory() as d:
        res = run_abundance(d)
    with open(os.path.join(HERE, "dory_subset_abund.json"), "w") as fp:
        json.dump(res, fp)
    print("wrote dory_subset_abund.json: %d unitigs, %d dominator rows, total abundance %d" % (
        len(res["cdbg"]), len(res["dom"]), sum(r[1] for r in res["cdbg"])))


if __name__ == "__main__":
    main()
