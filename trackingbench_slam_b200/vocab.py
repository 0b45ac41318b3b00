"""Host side of the bag-of-words row: the reference's ORBVocabulary (include/types/Frame.h:22, a
DBoW2::TemplatedVocabulary<FORB>) as flat arrays, its text format, and the BowVector / FeatureVector that
Frame::SetBow (src/types/Frame.cpp:267-270) derives from the per-feature transform.

The tree walk itself runs on the GPU (csrc/window.cu: bow_transform); nothing here computes distances.
"""
import numpy as np

# DBoW2 enums (third_part/DBoW2/DBoW2/BowVector.h)
TF_IDF, TF, IDF, BINARY = 0, 1, 2, 3
L1_NORM, L2_NORM, CHI_SQUARE, KL, BHATTACHARYYA, DOT_PRODUCT = 0, 1, 2, 3, 4, 5


class Vocabulary:
    """Nodes in id order (node 0 = root). parent[0] = -1; is_leaf = the file's leaf flag."""

    def __init__(self, k, L, parent, is_leaf, desc, weight, scoring=L1_NORM, weighting=TF_IDF):
        self.k, self.L, self.scoring, self.weighting = int(k), int(L), int(scoring), int(weighting)
        self.parent = np.ascontiguousarray(parent, np.int32)
        self.is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        self.desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        self.weight = np.ascontiguousarray(weight, np.float64)
        n = self.n_nodes
        assert self.is_leaf.shape == (n,) and self.desc.shape == (n, 32) and self.weight.shape == (n,)
        assert n >= 2 and self.parent[0] == -1 and np.all(self.parent[1:] < np.arange(1, n)) and np.all(self.parent[1:] >= 0)

    @property
    def n_nodes(self):
        return self.parent.shape[0]

    def csr(self):
        """(child_start, children, word_id): children in ascending id (loadFromTextFile order,
        TemplatedVocabulary.h:1393), word ids to flagged leaves in file order (:1403-1410)."""
        n = self.n_nodes
        cnt = np.bincount(self.parent[1:], minlength=n)
        start = np.zeros(n + 1, np.int32)
        start[1:] = np.cumsum(cnt)
        children = (np.argsort(self.parent[1:], kind='stable') + 1).astype(np.int32)
        word = np.full(n, -1, np.int32)
        flagged = self.is_leaf.astype(bool)
        flagged[0] = False
        word[flagged] = np.arange(int(flagged.sum()), dtype=np.int32)
        childless = (cnt == 0) & ~flagged
        childless[0] = False
        word[childless] = 0  # Node() default word_id (:315)
        return start, children, word

    def save_text(self, path):
        """ORBvoc.txt layout (loadFromTextFile :1338-1420): 'k L scoring weighting', then one line per node
        'parent isLeaf b0 .. b31 weight'. No trailing newline: the reference's `while(!f.eof())` loop would turn a
        final empty line into an extra root child with an uninitialised descriptor."""
        lines = ['%d %d %d %d' % (self.k, self.L, self.scoring, self.weighting)]
        for i in range(1, self.n_nodes):
            lines.append('%d %d %s %.17g' % (self.parent[i], self.is_leaf[i], ' '.join(str(int(b)) for b in self.desc[i]), self.weight[i]))
        with open(path, 'w') as f:
            f.write('\n'.join(lines))

    @staticmethod
    def load_text(path):
        """Parse the text vocabulary. Blank lines are skipped (see save_text for what the reference does with one)."""
        with open(path) as f:
            head = f.readline().split()
            k, L, scoring, weighting = (int(x) for x in head[:4])
            parent, leaf, desc, weight = [-1], [0], [np.zeros(32, np.uint8)], [0.0]
            for line in f:
                t = line.split()
                if not t:
                    continue
                parent.append(int(t[0])); leaf.append(1 if int(t[1]) > 0 else 0)
                desc.append(np.array([int(x) for x in t[2:34]], np.uint8)); weight.append(float(t[34]))
        return Vocabulary(k, L, parent, leaf, np.stack(desc), weight, scoring, weighting)


def synth_vocabulary(k=10, L=3, seed=0, flip=40, stop_frac=0.02):
    """Complete k-ary tree of depth L with random 256-bit node descriptors: a child = its parent with `flip` random
    bits toggled, so features have a meaningful nearest branch. Nodes in breadth-first order; a few words get
    weight 0 (stopped words, exercising the `w > 0` filter of transform())."""
    rng = np.random.Generator(np.random.PCG64(seed))
    parent, leaf, desc, weight = [-1], [0], [np.zeros(32, np.uint8)], [0.0]
    frontier = [0]
    root_desc = rng.integers(0, 256, 32, dtype=np.uint8)
    for level in range(1, L + 1):
        nxt = []
        for p in frontier:
            base = root_desc if p == 0 else desc[p]
            for _ in range(k):
                bits = np.unpackbits(base)
                idx = rng.choice(256, size=(128 if p == 0 else flip), replace=False)
                bits[idx] ^= 1
                parent.append(p); leaf.append(1 if level == L else 0)
                desc.append(np.packbits(bits))
                w = float(rng.uniform(0.5, 8.0)) if level == L else 0.0
                if level == L and rng.uniform() < stop_frac:
                    w = 0.0
                weight.append(w)
                nxt.append(len(parent) - 1)
        frontier = nxt
    return Vocabulary(k, L, parent, leaf, np.stack(desc), weight)


def bow_vector(word_id, weight, weighting=TF_IDF, scoring=L1_NORM):
    """BowVector of transform(features, v, fv, levelsup) (TemplatedVocabulary.h:1127-1194) from the per-feature
    (word, weight): dict-ordered (ids ascending, values float64). TF/TF_IDF accumulate weights in feature order
    (BowVector::addWeight), IDF/BINARY keep the first; L1/L2 scoring normalises (BowVector::normalize), the others
    divide TF weights by the number of words."""
    word_id = np.asarray(word_id); weight = np.asarray(weight, np.float64)
    acc = {}
    for w, v in zip(word_id.tolist(), weight.tolist()):
        if not v > 0:
            continue
        if weighting in (TF, TF_IDF):
            acc[w] = acc.get(w, 0.0) + v
        else:
            acc.setdefault(w, v)
    ids = np.array(sorted(acc), np.int32)
    vals = np.array([acc[i] for i in ids.tolist()], np.float64)
    must = scoring in (L1_NORM, L2_NORM)
    if weighting in (TF, TF_IDF) and len(ids) and not must:
        vals = vals / float(len(ids))
    if must and len(ids):
        norm = 0.0
        if scoring == L1_NORM:
            for v in vals.tolist():
                norm += abs(v)
        else:
            for v in vals.tolist():
                norm += v * v
            norm = float(np.sqrt(norm))
        if norm > 0.0:
            vals = vals / norm
    return ids, vals


def feature_vector(node_id, weight):
    """FeatureVector (std::map<NodeId, vector<feature idx>>) as per-feature node ids: -1 where the word is stopped."""
    node_id = np.asarray(node_id, np.int32).copy()
    node_id[~(np.asarray(weight) > 0)] = -1
    return node_id


def synth_vocabulary_fast(k=10, L=6, seed=0, p_flip=0.12):
    """ORBvoc-sized (k=10, L=6: 1.1 M nodes) synthetic tree, vectorised: a child = its parent with each bit toggled
    with probability p_flip (root children: 0.5). Breadth-first node order, every leaf at depth L, weights in [0.5, 8)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    parent = [np.array([-1], np.int32)]
    desc = [np.zeros((1, 32), np.uint8)]
    first, count = 0, 1
    prev = rng.integers(0, 256, (1, 32), dtype=np.uint8)
    for level in range(1, L + 1):
        n = count * k
        par = np.repeat(np.arange(first, first + count, dtype=np.int32), k)
        d = np.repeat(prev, k, axis=0)
        for s in range(0, n, 1 << 16):
            e = min(n, s + (1 << 16))
            d[s:e] ^= np.packbits(rng.random((e - s, 256)) < (0.5 if level == 1 else p_flip), axis=1)
        parent.append(par); desc.append(d)
        first, count, prev = first + count, n, d
    parent = np.concatenate(parent); desc = np.concatenate(desc)
    n_nodes = len(parent)
    is_leaf = np.zeros(n_nodes, np.uint8); is_leaf[first:] = 1
    weight = np.zeros(n_nodes, np.float64); weight[first:] = rng.uniform(0.5, 8.0, count)
    return Vocabulary(k, L, parent, is_leaf, desc, weight)
