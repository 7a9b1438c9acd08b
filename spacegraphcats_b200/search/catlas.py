"""Search-side CAtlas: the loaders and traversals the hot path consumes (reference:
``spacegraphcats/search/catlas.py`` -- catlas.csv / first_doms.txt loading :43-99, child-before-
parent iteration :142-154, ``leaves`` / ``shadow`` :174-211).  Host-side and tiny; kept because
per-dominator counting (K6) takes its level structure and because ``shadow`` output is a parity
target ("extracted neighbourhood node sets").
"""
import os
from collections import defaultdict


class CAtlas:
    def __init__(self, cdbg_directory, catlas_directory, load_domfile=True, load_sizefile=False, min_abund=0.0):
        self.cdbg_dir = cdbg_directory
        self.name = catlas_directory
        self.parent = {}
        self.children = defaultdict(set)
        self.levels = {}
        self._cdbg_to_catlas = {}
        self.max_level = -1
        self.root = -1
        self._read_catlas_csv(os.path.join(catlas_directory, "catlas.csv"))
        if load_domfile:
            self._read_first_doms(os.path.join(catlas_directory, "first_doms.txt"))
        if load_sizefile:
            self._read_size_info(os.path.join(cdbg_directory, "contigs.info.csv"), min_abund)

    # one row per node: node_id,cdbg_id,level,"child child ..."
    def _read_catlas_csv(self, path):
        with open(path, "rt") as fp:
            for row in fp:
                node, cdbg, level, kids = row.strip().split(",")
                node, level = int(node), int(level)
                kids = kids.split()
                if kids:
                    kid_ids = {int(x) for x in kids}
                    self.children[node] = kid_ids
                    for c in kid_ids:
                        self.parent[c] = node
                self.levels[node] = level
                if level > self.max_level:
                    self.max_level, self.root = level, node
                if level == 1:
                    self._cdbg_to_catlas[int(cdbg)] = node

    # one row per level-1 dominator: dom_cdbg_id member member ...
    def _read_first_doms(self, path):
        self.layer1_to_cdbg = {}
        self.cdbg_to_layer1 = {}
        with open(path, "rt") as fp:
            for row in fp:
                fields = row.split()
                if not fields:
                    continue
                node = self._cdbg_to_catlas[int(fields[0])]
                beneath = {int(x) for x in fields[1:]}
                self.layer1_to_cdbg[node] = beneath
                for c in beneath:
                    self.cdbg_to_layer1[c] = node

    def _read_size_info(self, path, min_abund):
        import csv

        self.cdbg_sizes = defaultdict(int)
        self.weighted_cdbg_sizes = defaultdict(float)
        with open(path, "rt") as fp:
            for row in csv.DictReader(fp):
                abund = float(row["mean_abund"])
                if not min_abund or abund >= min_abund:
                    cid, nk = int(row["contig_id"]), int(row["n_kmers"])
                    self.cdbg_sizes[cid] = nk
                    self.weighted_cdbg_sizes[cid] = abund * nk
        self.kmer_sizes, self.weighted_kmer_sizes = {}, {}
        for v in self:
            if self.levels[v] == 1:
                below = self.layer1_to_cdbg.get(v)
                self.kmer_sizes[v] = sum(self.cdbg_sizes[c] for c in below)
                self.weighted_kmer_sizes[v] = sum(self.weighted_cdbg_sizes[c] for c in below)
            else:
                self.kmer_sizes[v] = sum(self.kmer_sizes[c] for c in self.children[v])
                self.weighted_kmer_sizes[v] = sum(self.weighted_kmer_sizes[c] for c in self.children[v])

    def __iter__(self):
        """Every node after all of its children (breadth-first order from the root, reversed)."""
        order = [self.root]
        i = 0
        while i < len(self.levels):
            order.extend(self.children[order[i]])
            i += 1
        return reversed(order)

    def __len__(self):
        return len(self.parent)

    def decorate_with_shadow_sizes(self):
        self.shadow_sizes = {}
        for v in self:
            if self.levels[v] == 1:
                self.shadow_sizes[v] = len(self.layer1_to_cdbg[v])
            else:
                self.shadow_sizes[v] = sum(self.shadow_sizes[c] for c in self.children[v])

    def decorate_with_index_sizes(self, index):
        self.index_sizes = index.build_catlas_node_sizes(self)

    def leaves(self, nodes=None):
        """Leaf descendants of ``nodes`` (default: of the root)."""
        stack = list([self.root] if nodes is None else nodes)
        seen, out = set(), set()
        while stack:
            v = stack.pop()
            if v in seen:
                continue
            seen.add(v)
            kids = self.children[v]
            if kids:
                stack.extend(kids)
            else:
                out.add(v)
        return out

    def shadow(self, nodes):
        """cDBG vertices below the given catlas nodes."""
        out = set()
        for leaf in self.leaves(nodes):
            out.update(self.layer1_to_cdbg[leaf])
        return out
