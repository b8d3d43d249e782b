"""PedigreeDAG with the reference's call surface (pedigree/pedigree_dag.py) on flat index arrays.

The reference keeps a networkx.DiGraph plus six dict/set indexes and re-extracts a sub-DAG per animal.
Here the pedigree is a handful of numpy arrays (ids, sire/dam index, sex, generation, children CSR)
built once; the dict/set attributes of the reference (`males`, `females`, `kid2sire`, `kid2dam`,
`sire2kid`, `dam2kid`, `generationIndex`) are materialised lazily for callers that use them.
The arrays are what `genofix_b200.engine` hands to the C ABI (gfx_pedigree_desc).
"""
from collections import defaultdict

import numpy as np


class PedigreeDAG(object):
    def __init__(self, incoming_graph_data=None):
        # reference: PedigreeDAG(other) copies the indexes (pedigree_dag.py:16-33)
        if incoming_graph_data is not None:
            o = incoming_graph_data
            self._set_arrays(o.ids.copy(), o.sire_idx.copy(), o.dam_idx.copy(), o.sex.copy(),
                             o._edge_kid_rank.copy())
        else:
            self._set_arrays(np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int32),
                             np.zeros(0, np.int8), np.zeros(0, np.int64))

    # ------------------------------------------------------------------------------------------
    # construction
    # ------------------------------------------------------------------------------------------
    def _set_arrays(self, ids, sire_idx, dam_idx, sex, edge_kid_rank):
        self.ids = ids                    # sorted unique animal ids
        self.sire_idx = sire_idx          # index into ids or -1
        self.dam_idx = dam_idx
        self.sex = sex                    # 1 male, 2 female, 0 neither
        self._edge_kid_rank = edge_kid_rank  # first table row of each animal as a kid (child order)
        n = len(ids)
        self._index = None
        self._lazy = {}
        self._both_sexes = np.zeros(n, bool)
        # generation = min(pure sire-line length, pure dam-line length)  (pedigree_dag.py:171-175,197-200)
        self.generation = np.minimum(self._line_depth(sire_idx), self._line_depth(dam_idx)).astype(np.int32) \
            if n else np.zeros(0, np.int32)
        # children CSR in edge-insertion order = order of the child's first row in the table
        if n:
            kid = np.concatenate([np.nonzero(sire_idx >= 0)[0], np.nonzero(dam_idx >= 0)[0]])
            par = np.concatenate([sire_idx[sire_idx >= 0], dam_idx[dam_idx >= 0]])
            order = np.lexsort((edge_kid_rank[kid], par))
            self._kid_sorted = kid[order].astype(np.int32)
            counts = np.bincount(par, minlength=n)
            self._kid_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        else:
            self._kid_sorted = np.zeros(0, np.int32)
            self._kid_off = np.zeros(1, np.int32)

    @staticmethod
    def _line_depth(parent_idx):
        n = len(parent_idx)
        depth = np.zeros(n, dtype=np.int64)
        cur = np.arange(n)
        alive = np.ones(n, dtype=bool)
        for _ in range(n + 1):
            nxt = parent_idx[cur]
            alive &= nxt >= 0
            if not alive.any():
                break
            depth[alive] += 1
            cur = np.where(alive, nxt, cur)
        return depth

    @staticmethod
    def from_table(pedigree):
        """rows of (kid, sire, dam, sex); ids <= 0 mean unknown (pedigree_dag.py:188-214)."""
        t = np.asarray(pedigree)
        if t.ndim != 2 or t.shape[1] < 4:
            raise ValueError("pedigree table needs columns kid, sire, dam, sex")
        t = t[:, :4].astype(np.int64)
        kid, sire, dam, sx = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
        ids = np.unique(np.concatenate([kid[kid > 0], sire[sire > 0], dam[dam > 0]]))
        n = len(ids)
        pos = lambda a: np.searchsorted(ids, a)  # noqa: E731
        sire_idx = np.full(n, -1, np.int32)
        dam_idx = np.full(n, -1, np.int32)
        rank = np.full(n, np.iinfo(np.int64).max, np.int64)
        ok = kid > 0
        rows = np.nonzero(ok)[0]
        ki = pos(kid[ok])
        # later rows win, like the dict comprehension of the reference
        m = sire[ok] > 0
        sire_idx[ki[m]] = pos(sire[ok][m])
        m = dam[ok] > 0
        dam_idx[ki[m]] = pos(dam[ok][m])
        # males = sires + kids with sex 1, females = dams + kids with sex 2 (pedigree_dag.py:190-191)
        is_male = np.zeros(n, bool)
        is_female = np.zeros(n, bool)
        is_male[ki[sx[ok] == 1]] = True
        is_female[ki[sx[ok] == 2]] = True
        is_male[pos(sire[sire > 0])] = True
        is_female[pos(dam[dam > 0])] = True
        sex = np.where(is_male, 1, np.where(is_female, 2, 0)).astype(np.int8)
        np.minimum.at(rank, ki, rows)
        dag = PedigreeDAG()
        dag._set_arrays(ids, sire_idx, dam_idx, sex, rank)
        dag._both_sexes = is_male & is_female
        dag._table = t      # kept so that `males` / `females` can be built in the reference's insertion order
        return dag

    @staticmethod
    def from_file(file, delimiter=' '):
        pedigree = np.loadtxt(file, delimiter=delimiter, dtype=np.int64, usecols=(0, 1, 2, 3), ndmin=2)
        return PedigreeDAG.from_table(pedigree)

    # ------------------------------------------------------------------------------------------
    # index helpers
    # ------------------------------------------------------------------------------------------
    def index_of(self, animal_ids):
        """Vectorised id -> pedigree index (-1 if absent)."""
        a = np.asarray(animal_ids, dtype=np.int64)
        p = np.searchsorted(self.ids, a)
        p = np.minimum(p, max(len(self.ids) - 1, 0))
        ok = (len(self.ids) > 0) & (self.ids[p] == a) if len(self.ids) else np.zeros(a.shape, bool)
        return np.where(ok, p, -1).astype(np.int64)

    def _idx(self, animal):
        if self._index is None:
            self._index = {int(k): i for i, k in enumerate(self.ids)}
        return self._index.get(int(animal), -1)

    def _children_idx(self, i):
        return self._kid_sorted[self._kid_off[i]:self._kid_off[i + 1]]

    def children_csr_reference_order(self):
        """Children CSR whose per-parent order is `set(list(get_kids(parent)))` iteration order, the
        order the reference feeds children to generate_probs_kids (correct_genotypes3.py:199-200,
        256-258).  Only the float rounding of the children term depends on it."""
        off = self._kid_off
        out = np.empty(len(self._kid_sorted), dtype=np.int32)
        ids = self.ids
        pos = self._idx
        for p in np.nonzero(np.diff(off) > 1)[0]:
            kids = [int(ids[k]) for k in self._kid_sorted[off[p]:off[p + 1]]]
            out[off[p]:off[p + 1]] = [pos(k) for k in set(kids)]
        single = np.nonzero(np.diff(off) == 1)[0]
        out[off[single]] = self._kid_sorted[off[single]]
        return off.copy(), out

    # ------------------------------------------------------------------------------------------
    # the reference's dict / set attributes, built on demand
    # ------------------------------------------------------------------------------------------
    def _get(self, name, build):
        if name not in self._lazy:
            self._lazy[name] = build()
        return self._lazy[name]

    @property
    def males(self):
        def build():
            t = getattr(self, "_table", None)
            if t is None:
                return set(int(x) for x in self.ids[self.sex == 1])
            # the reference's own expression (pedigree_dag.py:190): same elements, same insertion order, so that
            # iterating the set (split_pedigree, generationIndex) visits the animals in the reference's order
            return set([int(r[1]) for r in t if int(r[1]) > 0] + [int(r[0]) for r in t if r[3] == 1 and int(r[0]) > 0])
        return self._get("males", build)

    @property
    def females(self):
        def build():
            t = getattr(self, "_table", None)
            if t is None:
                return set(int(x) for x in self.ids[(self.sex == 2) | self._both_sexes])
            return set([int(r[2]) for r in t if int(r[2]) > 0] + [int(r[0]) for r in t if r[3] == 2 and int(r[0]) > 0])   # :191
        return self._get("females", build)

    @property
    def kid2sire(self):
        return self._get("kid2sire", lambda: {int(self.ids[k]): int(self.ids[self.sire_idx[k]])
                                              for k in np.nonzero(self.sire_idx >= 0)[0]})

    @property
    def kid2dam(self):
        return self._get("kid2dam", lambda: {int(self.ids[k]): int(self.ids[self.dam_idx[k]])
                                             for k in np.nonzero(self.dam_idx >= 0)[0]})

    def _one2many(self, parent_idx):
        d = defaultdict(set)
        for k in np.nonzero(parent_idx >= 0)[0]:
            d[int(self.ids[parent_idx[k]])].add(int(self.ids[k]))
        return d

    @property
    def sire2kid(self):
        return self._get("sire2kid", lambda: self._one2many(self.sire_idx))

    @property
    def dam2kid(self):
        return self._get("dam2kid", lambda: self._one2many(self.dam_idx))

    @property
    def generationIndex(self):
        def build():
            d = defaultdict(set)
            sexed = self.sex > 0
            for i in np.nonzero(sexed)[0]:
                d[int(self.generation[i])].add(int(self.ids[i]))
            return d
        return self._get("generationIndex", build)

    # ------------------------------------------------------------------------------------------
    # graph-like surface used by callers of the reference class
    # ------------------------------------------------------------------------------------------
    def __contains__(self, animal):
        try:
            return self._idx(animal) >= 0
        except (TypeError, ValueError):
            return False

    def __len__(self):
        return len(self.ids)

    def has_node(self, animal):
        return animal in self

    @property
    def nodes(self):
        return [int(x) for x in self.ids]

    def get_nodes(self):
        return list(self.males) + list(self.females)

    def _require(self, animal):
        i = self._idx(animal)
        if i < 0:
            try:
                import networkx as nx
                raise nx.NetworkXError("The node %s is not in the digraph." % (animal,))
            except ImportError:
                raise KeyError(animal)
        return i

    def successors(self, parent):
        i = self._require(parent)
        return iter([int(self.ids[k]) for k in self._children_idx(i)])

    def predecessors(self, kid):
        i = self._require(kid)
        return iter([int(self.ids[p]) for p in (self.sire_idx[i], self.dam_idx[i]) if p >= 0])

    def get_parents(self, kid):
        """(sire, dam) ids or None (pedigree_dag.py:245-268)."""
        i = self._idx(kid)
        if i < 0:
            return (None, None)
        s, d = self.sire_idx[i], self.dam_idx[i]
        return (int(self.ids[s]) if s >= 0 else None, int(self.ids[d]) if d >= 0 else None)

    def get_kids(self, parent):
        return self.successors(parent)

    def get_parents_depth(self, kids, depth):
        """Ancestors level by level, duplicates kept (pedigree_dag.py:221-231)."""
        if isinstance(kids, (int, np.integer)):
            kids = [kids]
        kids = list(kids)
        if depth > 0 and len(kids) > 0:
            new = [p for kid in kids for p in self.get_parents(kid) if p is not None]
            yield from new
            yield from self.get_parents_depth(new, depth - 1)

    def get_kids_depth(self, kids, depth):
        if isinstance(kids, (int, np.integer)):
            kids = [kids]
        kids = list(kids)
        if depth > 0 and len(kids) > 0:
            new = [p for kid in kids for p in self.get_kids(kid)]
            yield from new
            yield from self.get_kids_depth(new, depth - 1)

    def balance_nodes(self, nodes):
        """Close the node list under 'partner of an included parent' (pedigree_dag.py:124-139).
        KeyError for a node with exactly one known parent, like the reference."""
        nodes = list(nodes)
        inset = set(nodes)
        changed = True
        while changed:
            changed = False
            for node in list(nodes):
                i = self._idx(node)
                if i < 0 or (self.sire_idx[i] < 0 and self.dam_idx[i] < 0):
                    continue
                if self.sire_idx[i] < 0 or self.dam_idx[i] < 0:
                    raise KeyError(node)
                s, d = int(self.ids[self.sire_idx[i]]), int(self.ids[self.dam_idx[i]])
                if s in inset and d not in inset:
                    nodes.append(d)
                    inset.add(d)
                    changed = True
                if d in inset and s not in inset:
                    nodes.append(s)
                    inset.add(s)
                    changed = True
        return nodes

    def get_subset(self, nodes, balance_parents=True):
        """Induced sub-pedigree (pedigree_dag.py:141-168)."""
        if balance_parents:
            nodes = self.balance_nodes(nodes)
        keep = set(int(x) for x in nodes)
        rows = []
        for k in sorted(x for x in keep if x in self):
            i = self._idx(k)
            s, d = self.get_parents(k)
            rows.append((k, s if s in keep else 0, d if d in keep else 0, int(self.sex[i])))
        if not rows:
            return PedigreeDAG()
        sub = PedigreeDAG.from_table(np.array(rows, dtype=np.int64))
        # keep the sex of animals that lost their role as a parent in the subset
        for k, _s, _d, sx in rows:
            sub.sex[sub._idx(k)] = sx
        sub._lazy = {}
        return sub

    @staticmethod
    def recursiveDepth(current, dictionary, count=0):
        while current in dictionary:
            current = dictionary[current]
            count += 1
        return count

    def get_all_siblings(self, kid):
        sire, dam = self.get_parents(kid)
        sibs = set()
        if sire is not None:
            sibs.update(self.get_kids(sire))
        if dam is not None:
            sibs.update(self.get_kids(dam))
        sibs.discard(kid)
        return list(sibs)

    def get_full_siblings(self, kid):
        sire, dam = self.get_parents(kid)
        sibs = set()
        if sire is not None and dam is not None:
            sibs = set(self.get_kids(sire)).intersection(set(self.get_kids(dam)))
        sibs.discard(kid)
        return list(sibs)

    def _reach(self, start, nxt):
        seen = set()
        stack = [start]
        while stack:
            x = stack.pop()
            for y in nxt(x):
                if y not in seen:
                    seen.add(y)
                    stack.append(y)
        return seen

    def get_ancestors(self, parent):
        return self._reach(parent, lambda x: [p for p in self.get_parents(x) if p is not None])

    def get_decendents(self, parent):
        return self._reach(parent, lambda x: list(self.get_kids(x)))

    def get_partners(self, parent):
        partners = set()
        for kid in self.get_kids(parent):
            partners.update(p for p in self.get_parents(kid) if p is not None)
        partners.discard(parent)
        return list(partners)

    def get_grandkids(self, parent):
        return [g for kid in self.get_kids(parent) for g in self.get_kids(kid)]

    def get_relationship(self, kid, other):
        if other in list(self.get_kids(kid)):
            return "son" if other in self.males else "daughter"
        elif other in self.get_parents(kid):
            return "sire" if other in self.males else "dam"
        elif other in self.get_grandkids(kid):
            return "grandson" if other in self.males else "granddaughter"
        elif kid in self.get_grandkids(other):
            return "grandfather" if other in self.males else "grandmother"
        elif other in self.get_partners(kid):
            return "husband" if other in self.males else "wife"
        return "other"

    def as_ped(self):
        """(kid, sire, dam, sex) for animals with both parents, generation by generation
        (pedigree_dag.py:360-369)."""
        lines = []
        both = (self.sire_idx >= 0) & (self.dam_idx >= 0) & (self.sex > 0)
        for g in sorted(set(int(x) for x in self.generation[both])):
            for i in np.nonzero(both & (self.generation == g))[0]:
                lines.append((int(self.ids[i]), int(self.ids[self.sire_idx[i]]), int(self.ids[self.dam_idx[i]]),
                              1 if self.sex[i] == 1 else 2))
        return lines

    def split_pedigree(self, min_cluster_size=100):
        """Clusters of animals around connected mating pairs (pedigree_dag.py:41-122).  The reference's procedure is
        order dependent (first matching cluster wins, clusters are scanned by size, Python set iteration order decides
        ties), so it is followed step by step on the same Python sets; only the membership tests use an index instead
        of scanning.  Returns a list of sets of ids like the reference."""
        allkids = self.males.union(self.females)
        partner_pairs = list(set([(sire, dam) for sire, dam in [self.get_parents(kid) for kid in allkids]
                                  if sire is not None and dam is not None]))                       # :42
        clustered_pairs = []
        member = defaultdict(list)      # animal -> positions of the clusters holding it (ascending; valid until the merges)

        def first_cluster(a, b):
            """position of the first cluster that holds a or b, and which of the two it holds first (a wins a tie)"""
            ia = member[a][0] if a is not None and member.get(a) else None
            ib = member[b][0] if b is not None and member.get(b) else None
            if ia is None and ib is None:
                return None, None
            if ib is None or (ia is not None and ia <= ib):
                return ia, 0
            return ib, 1
        for sire, dam in partner_pairs:                                                               # :44-56
            pos, which = first_cluster(sire, dam)
            if pos is None:
                clustered_pairs.append({sire, dam})
                member[sire].append(len(clustered_pairs) - 1)
                member[dam].append(len(clustered_pairs) - 1)
            else:
                new = dam if which == 0 else sire
                if new not in clustered_pairs[pos]:
                    clustered_pairs[pos].add(new)
                    lst = member[new]
                    lst.append(pos)
                    lst.sort()
        pairedkids = set([x[0] for x in partner_pairs] + [x[1] for x in partner_pairs])
        singlekids = [x for x in allkids if x not in pairedkids]
        for kid in singlekids:                                                                        # :61-74
            sire, dam = self.get_parents(kid)
            pos, _which = first_cluster(sire, dam)
            if pos is None:
                raise Exception("%s illegal lone node in pedigree" % kid)
            if kid not in clustered_pairs[pos]:
                clustered_pairs[pos].add(kid)
                lst = member[kid]
                lst.append(pos)
                lst.sort()
        # clusters that share an animal are merged, smallest first (:78-90); list positions change from here on
        for _i in range(len(clustered_pairs)):
            merged_any = False
            for cluster in sorted(clustered_pairs, key=lambda x: len(x)):
                for target in sorted([x for x in clustered_pairs if x != cluster], key=lambda x: len(x)):
                    if not cluster.isdisjoint(target):
                        target.update(cluster)
                        clustered_pairs.remove(cluster)
                        merged_any = True
                        break
            if not merged_any:
                break       # a pass without a merge leaves the list as it is: the remaining passes are no-ops
        k2s, k2d, s2k, d2k = self.kid2sire, self.kid2dam, self.sire2kid, self.dam2kid
        for size in range(min_cluster_size):                                                          # :94-121
            for _i in range(len(clustered_pairs)):
                found = False
                for cluster in [x for x in clustered_pairs if len(x) == size]:
                    sires = set([k2s[x] for x in cluster if x in k2s])
                    dams = set([k2d[x] for x in cluster if x in k2d])
                    parents = sires.union(dams)
                    for target in sorted([x for x in clustered_pairs if x != cluster], key=lambda x: len(x)):
                        if len(parents.intersection(target)) > 0:
                            target.update(cluster)
                            clustered_pairs.remove(cluster)
                            found = True
                            break
                    if found:
                        break
                    kids = [list(s2k[x]) for x in cluster if x in s2k]
                    if len(kids) > 0:
                        kids = set(np.concatenate(kids))
                    kids2 = [list(d2k[x]) for x in cluster if x in d2k]
                    if len(kids2) > 0:
                        kids = kids.union(set(np.concatenate(kids2)))      # AttributeError like the reference when `kids` is still a list
                    for target in sorted([x for x in clustered_pairs if x != cluster], key=lambda x: len(x)):
                        if len(kids.intersection(target)) > 0:
                            target.update(cluster)
                            clustered_pairs.remove(cluster)
                            found = True
                            break
                    if found:
                        break
                if not found:
                    break   # no cluster of this size can be merged any more: the remaining passes repeat the same scan
        return clustered_pairs
