"""Sequential restatement of ChaNGa's interaction-list walk, bucket after bucket,
with the per-level state that persists between buckets -- the way the reference
runs it.  TEST INFRASTRUCTURE ONLY (pure Python: small trees).

Follows, step by step:
  TreePiece::nextBucket / startNextBucket      TreePiece.cpp:3983-4072, 3702-3790
  TreePiece::getStartAncestor                  TreePiece.cpp:4757-4775
  LocalTargetWalk::walk / dft / processNode    TreeWalk.cpp:297-428
  ListCompute::doWork (Local opt)              Compute.cpp:690-884, Opt.h:86-128
  ListCompute::stateReady, GPU branch          Compute.cpp:1608-1863
  openCriterionNode / openSoftening            gravity.h:652-723, 251-260 (C: gravity_oracle.c)

The tree is given as flat arrays (children, particle ranges, tight boxes, 27-value
moment records in double); node ids are the breadth-first nodeArrayIndex.  The
product's parallel recursive walk (changa_b200/csrc/treewalk.cpp) must reproduce
these lists entry for entry, low offsetID bits included.
"""
import numpy as np

from . import oracle as orc

OFFSET_MASK = 0x1FF << 22


def encode_offset(req, x, y, z):
    return req | (((x + 3) | ((y + 3) << 3) | ((z + 3) << 6)) << 22)


def decode_offset(off, period):
    return np.array([((off >> 22) & 7) - 3, ((off >> 25) & 7) - 3, ((off >> 28) & 7) - 3], dtype=np.float64) * period


class SequentialWalk:
    def __init__(self, child0, child1, first, last, boxlo, boxhi, moments, bucket_node, theta, n_replicas,
                 period, bucket_active=None):
        self.c0, self.c1, self.first, self.last = child0, child1, first, last
        self.boxlo, self.boxhi = np.ascontiguousarray(boxlo), np.ascontiguousarray(boxhi)
        self.mom = np.ascontiguousarray(moments, dtype=np.float64)
        self.bucket_node = bucket_node
        self.theta, self.theta_mono = theta, theta ** 4
        self.nrep, self.period = n_replicas, period
        nn = len(child0)
        self.parent = np.full(nn, -1)
        self.level = np.zeros(nn, dtype=np.int64)
        for i in range(nn):
            for c in (child0[i], child1[i]):
                if c >= 0:
                    self.parent[c] = i
                    self.level[c] = self.level[i] + 1
        nb = len(bucket_node)
        self.active = np.ones(nb, dtype=bool) if bucket_active is None else np.asarray(bucket_active, dtype=bool)
        # buckets beneath each node (startBucket, numBucketsBeneath)
        self.bstart = np.zeros(nn, dtype=np.int64)
        self.bcount = np.zeros(nn, dtype=np.int64)
        for b, node in enumerate(bucket_node):
            self.bstart[node], self.bcount[node] = b, 1
        for i in range(nn - 1, -1, -1):
            if child0[i] < 0 and child1[i] < 0:
                continue
            kids = [c for c in (child0[i], child1[i]) if c >= 0]
            self.bstart[i] = self.bstart[kids[0]]
            self.bcount[i] = sum(self.bcount[c] for c in kids)
        nl = int(self.level.max()) + 2
        self.chk = [[] for _ in range(nl)]
        self.und = [[] for _ in range(nl)]
        self.clist = [[] for _ in range(nl)]
        self.lplist = [[] for _ in range(nl)]
        self.lowest = None

    def is_bucket(self, i):
        return self.c0[i] < 0 and self.c1[i] < 0

    def open_criterion(self, node, my, shift):
        return orc.lib().orc_open_criterion_node(
            self.mom[node], int(self.last[node] - self.first[node] + 1), shift, self.mom[my],
            self.boxlo[my], self.boxhi[my], int(self.is_bucket(my)), self.theta, self.theta_mono)

    def open_softening(self, node, my, shift):
        return orc.lib().orc_open_softening(self.mom[node], shift, self.mom[my], self.boxlo[my], self.boxhi[my])

    # ListCompute::doWork; returns True for KEEP
    def do_work(self, node, req, my, level):
        shift = decode_offset(req, self.period)
        open_ = self.open_criterion(node, my, shift)
        if open_ != 0:
            if self.is_bucket(node):                       # KEEP_LOCAL_BUCKET
                self.lplist[level].append((node, req))
                return False
            if open_ == 1 or self.is_bucket(my):           # CONTAIN, or INTERSECT under a bucket
                for c in (self.c0[node], self.c1[node]):
                    if c >= 0:
                        self.chk[level].append((c, req))
                return False
            self.und[level].append((node, req))            # INTERSECT
            return True
        self.clist[level].append((node, req))              # COMPUTE
        return False

    def dft(self, my, target_bucket, is_root, level):
        target_node = self.bucket_node[target_bucket]
        if not is_root:
            self.und[level], self.clist[level], self.lplist[level] = [], [], []   # initState
            assert not self.chk[level]
            for node, off in self.und[level - 1]:
                self.do_work(node, (target_bucket & ~OFFSET_MASK) | (off & OFFSET_MASK), my, level)
        while self.chk[level]:
            node, off = self.chk[level].pop(0)
            self.do_work(node, (target_bucket & ~OFFSET_MASK) | (off & OFFSET_MASK), my, level)
        if self.und[level]:
            # whichChild(targetKey): the child whose bucket range holds the target
            child = None
            for c in (self.c0[my], self.c1[my]):
                if c >= 0 and self.bstart[c] <= target_bucket < self.bstart[c] + self.bcount[c]:
                    child = c
            assert child is not None
            self.dft(child, target_bucket, False, level + 1)
        else:
            self.lowest = my

    def start_ancestor(self, current, previous):
        if previous < 0:
            return 0
        a, b = self.bucket_node[current], self.bucket_node[previous]
        pa = set()
        x = b
        while x >= 0:
            pa.add(x)
            x = self.parent[x]
        x, below = a, a
        while x not in pa:
            below = x
            x = self.parent[x]
        return below          # child of the LCA that holds `current`

    def run(self):
        """{bucket: (cells [(node, offsetID)], part buckets [(first, off, n)], softened [(node, offsetID)])}"""
        nb = len(self.bucket_node)
        out = {}
        cur, prev, placed = 0, -1, False
        while cur < nb:
            if not self.active[cur]:
                cur += 1
                continue
            anc = self.start_ancestor(cur, prev)
            lvl = int(self.level[anc])
            if not placed:
                placed = True
                r = self.nrep
                for x in range(-r, r + 1):
                    for y in range(-r, r + 1):
                        for z in range(-r, r + 1):
                            self.chk[lvl].append((0, encode_offset(0, x, y, z)))
            self.dft(anc, cur, lvl == 0, lvl)
            low = self.lowest
            maxlevel = int(self.level[low])
            end = int(self.bstart[low] + self.bcount[low])
            for b in range(cur, end):
                if not self.active[b]:
                    continue
                bn = self.bucket_node[b]
                cells, soft, parts = [], [], []
                for level in range(maxlevel + 1):
                    for node, off in self.clist[level]:
                        if self.open_softening(node, bn, decode_offset(off, self.period)):
                            soft.append((node, off))
                        else:
                            cells.append((node, off))
                for level in range(maxlevel + 1):
                    for node, off in self.lplist[level]:
                        parts.append((int(self.first[node]), off & OFFSET_MASK,
                                      int(self.last[node] - self.first[node] + 1)))
                out[b] = (cells, parts, soft)
            prev, cur = cur, end
        return out
