"""SECOND, independent restatement of the Go tree's hot path -- TEST INFRASTRUCTURE, written from the Go sources only
(pkg/pronet/alias.go:10-106, pkg/pronet/pronet.go:90-109 / :195-333, pkg/pronet/optimizer.go:8-117,
internal/models/line/line.go:73-206, internal/models/bpr/bpr.go:61-131, internal/models/deepwalk/deepwalk.go:61-141,
internal/models/node2vec/node2vec.go:82-173, internal/models/cpr/cpr.go:127-282, internal/models/tpr/tpr.go:101-262),
without consulting oracle/smore_oracle.cpp: plain Python floats (IEEE doubles, one operation at a time, the Go statement
order), no numpy in the arithmetic. No Go toolchain exists in this image, so the Go-semantics kernels cannot be pinned
against the Go binaries; what this file buys is that two independently written restatements (this one and the C++ oracle)
and the CUDA kernels agree bit for bit on the fixtures of tests/golden/golden_go_v1.npz.

The only convention shared with the rest of the repo is the draw stream, which replaces Go's math/rand:
word k of Philox stream (seed, stream);  rng.Intn(n) / Int63n(n) := (k * n) >> 32;  rng.Float64() := k * 2^-32.
(One source of disagreement with a real Go build remains and is documented there: Go's math.Pow / math.Exp are its own
implementations and may differ from libm in the last bit.)"""
import math

import numpy as np

from tests.synth_ref import philox4x32_10

MONITOR, POWER_SAMPLE, SIGMOID_TABLE_SIZE, MAX_SIGMOID = 10000, 0.75, 1000, 8.0  # pronet.go:14-19


class Words:
    """Sequential reader of Philox stream (seed, stream): word n = lane n & 3 of block n >> 2."""

    def __init__(self, seed, stream):
        self.seed, self.stream, self.pos, self.buf, self.base = seed, stream, 0, [], 0

    def next(self):
        if self.pos >= self.base + len(self.buf):
            self.base = self.pos & ~3
            blocks = np.arange(self.base >> 2, (self.base >> 2) + 256, dtype=np.uint64)
            r = philox4x32_10(blocks & np.uint64(0xFFFFFFFF), blocks >> np.uint64(32), np.uint64(self.stream & 0xFFFFFFFF),
                              np.uint64(self.stream >> 32), self.seed & 0xFFFFFFFF, (self.seed >> 32) & 0xFFFFFFFF)
            self.buf = np.stack(r, axis=1).reshape(-1).tolist()
        k = self.buf[self.pos - self.base]
        self.pos += 1
        return int(k)

    def intn(self, n):
        return (self.next() * n) >> 32

    def float64(self):
        return self.next() * (1.0 / 4294967296.0)


def build_alias_method(distribution, power):
    """alias.go:10-90."""
    n = len(distribution)
    prob, alias = [0.0] * n, [0] * n
    total = 0.0
    norm = [0.0] * n
    for i in range(n):
        norm[i] = math.pow(distribution[i], power) if distribution[i] > 0 else 0.0
        total += norm[i]
    if total == 0:
        return [1.0] * n, list(range(n))
    for i in range(n):
        norm[i] = norm[i] * float(n) / total
    small, large = [], []
    for i in range(n):
        (small if norm[i] < 1.0 else large).append(i)
    while small and large:
        l = small.pop()
        g = large.pop()
        prob[l], alias[l] = norm[l], g
        norm[g] = norm[g] + norm[l] - 1.0
        (small if norm[g] < 1.0 else large).append(g)
    while large:
        g = large.pop()
        prob[g], alias[g] = 1.0, g
    while small:
        l = small.pop()
        prob[l], alias[l] = 1.0, l
    return prob, alias


class ProNet:
    def __init__(self, src, dst, weight, undirected):
        """LoadEdgeList + buildGraph (pronet.go:112-249) on an in-memory edge list; src / dst are vertex NAMES."""
        self.hash, self.keys = {}, []
        self.graph, self.edge = {}, {}
        self.max_line = 0
        for a, b, w in zip(src, dst, weight):
            v1, v2 = self._vid(a), self._vid(b)
            self.graph.setdefault(v1, []).append(v2)
            self.edge.setdefault(v1, []).append(float(w))
            if undirected:
                self.graph.setdefault(v2, []).append(v1)
                self.edge.setdefault(v2, []).append(float(w))
            self.max_line += 1
        self.max_vid = len(self.keys)
        self.out_degree = [0.0] * self.max_vid
        self.in_degree = [0.0] * self.max_vid  # Vertices[..].InDegree (Contexts[..].InDegree holds the same sums)
        for vid in range(self.max_vid):
            if vid in self.graph:
                for nid, w in zip(self.graph[vid], self.edge[vid]):
                    self.out_degree[vid] += w
                    self.in_degree[nid] += w
        self.vertex_at = build_alias_method(self.out_degree, 1.0)
        self.negative_at = build_alias_method([i + o for i, o in zip(self.in_degree, self.out_degree)], POWER_SAMPLE)
        self.sigmoid = [1.0 / (1.0 + math.exp(-(float(i) * 2.0 * MAX_SIGMOID / float(SIGMOID_TABLE_SIZE) - MAX_SIGMOID)))
                        for i in range(SIGMOID_TABLE_SIZE + 1)]

    def _vid(self, name):
        if name not in self.hash:
            self.hash[name] = len(self.keys)
            self.keys.append(name)
        return self.hash[name]

    def fast_sigmoid(self, x):  # pronet.go:98-109
        if x < -MAX_SIGMOID:
            return 0.0
        if x > MAX_SIGMOID:
            return 1.0
        idx = int((x + MAX_SIGMOID) * float(SIGMOID_TABLE_SIZE) / MAX_SIGMOID / 2.0)
        return self.sigmoid[min(idx, len(self.sigmoid) - 1)]

    @staticmethod
    def alias_sample(table, rng):  # alias.go:93-106
        prob, alias = table
        i = rng.intn(len(prob))
        r = rng.float64()
        return i if r < prob[i] else alias[i]

    def source_sample(self, rng):
        return self.alias_sample(self.vertex_at, rng)

    def negative_sample(self, rng):
        return self.alias_sample(self.negative_at, rng)

    def target_sample(self, vid, rng):  # pronet.go:257-284
        neighbors = self.graph.get(vid, [])
        if not neighbors:
            return -1
        weights = self.edge[vid]
        total = 0.0
        for w in weights:
            total += w
        r = rng.float64() * total
        cum = 0.0
        for i, w in enumerate(weights):
            cum += w
            if r <= cum:
                return neighbors[i]
        return neighbors[-1]

    def random_walk(self, vid, steps, rng):  # pronet.go:292-307
        walk = [vid]
        cur = vid
        for _ in range(steps):
            nxt = self.target_sample(cur, rng)
            if nxt == -1:
                break
            walk.append(nxt)
            cur = nxt
        return walk

    @staticmethod
    def skip_grams(walk, window):  # pronet.go:310-333
        vs, cs = [], []
        for i in range(len(walk)):
            for j in range(max(i - window, 0), min(i + window + 1, len(walk))):
                if i != j:
                    vs.append(walk[i])
                    cs.append(walk[j])
        return vs, cs

    # ---- optimizer.go ----
    def sgd_update(self, ve, ce, label, alpha, vgrad, cgrad):  # :61-84
        score = 0.0
        for d in range(len(ve)):
            score += ve[d] * ce[d]
        grad = alpha * (label - self.fast_sigmoid(score))
        for d in range(len(ve)):
            vgrad[d] += grad * ce[d]
            cgrad[d] += grad * ve[d]

    def update_pair(self, wv, wc, vertex, context, dim, K, alpha, rng):  # :21-58
        vgrad, cgrad = [0.0] * dim, [0.0] * dim
        self.sgd_update(wv[vertex], wc[context], 1.0, alpha, vgrad, cgrad)
        for _ in range(K):
            neg = self.negative_sample(rng)
            if neg == context:
                continue
            ngrad = [0.0] * dim
            self.sgd_update(wv[vertex], wc[neg], 0.0, alpha, vgrad, ngrad)
            for d in range(dim):
                wc[neg][d] += ngrad[d]
        for d in range(dim):
            wv[vertex][d] += vgrad[d]
            wc[context][d] += cgrad[d]

    def update_bpr_pair(self, wv, wc, vertex, pos, neg, dim, alpha, lam):  # :87-117
        ps = ns = 0.0
        for d in range(dim):
            ps += wv[vertex][d] * wc[pos][d]
            ns += wv[vertex][d] * wc[neg][d]
        coef = alpha * self.fast_sigmoid(ns - ps)
        for d in range(dim):
            vg = coef * (wc[pos][d] - wc[neg][d])
            pg = coef * wv[vertex][d]
            ng = -coef * wv[vertex][d]
            wv[vertex][d] += vg - lam * alpha * wv[vertex][d]
            wc[pos][d] += pg - lam * alpha * wc[pos][d]
            wc[neg][d] += ng - lam * alpha * wc[neg][d]


def update_first_order(pn, wv, source, target, dim, K, alpha, rng):  # line.go:153-200
    vgrad, cgrad = [0.0] * dim, [0.0] * dim
    score = 0.0
    for d in range(dim):
        score += wv[source][d] * wv[target][d]
    grad = alpha * (1.0 - pn.fast_sigmoid(score))
    for d in range(dim):
        vgrad[d] = grad * wv[target][d]
        cgrad[d] = grad * wv[source][d]
    for _ in range(K):
        neg = pn.negative_sample(rng)
        if neg == target or neg == source:
            continue
        score = 0.0
        for d in range(dim):
            score += wv[source][d] * wv[neg][d]
        grad = alpha * (0.0 - pn.fast_sigmoid(score))
        for d in range(dim):
            vgrad[d] += grad * wv[neg][d]
            wv[neg][d] += grad * wv[source][d]
    for d in range(dim):
        wv[source][d] += vgrad[d]
        wv[target][d] += cgrad[d]


def _schedule(count, total, alpha, cur):  # line.go:133-142 / bpr.go:118-127 / deepwalk.go:121-130
    if count % MONITOR == 0:
        cur = alpha * (1.0 - float(count) / float(total))
        if cur < alpha * 0.0001:
            cur = alpha * 0.0001
    return cur


def train_line(pn, wv, wc, order, dim, iterations, total, K, alpha, rng):
    """LINE.Train (line.go:73-150), one worker, `iterations` trips of the sample loop of a run of `total` samples."""
    cur, count = alpha, 0
    for _ in range(iterations):
        source = pn.source_sample(rng)
        target = pn.target_sample(source, rng)
        if target == -1:
            continue
        if order == 1:
            update_first_order(pn, wv, source, target, dim, K, cur, rng)
        else:
            pn.update_pair(wv, wc, source, target, dim, K, cur, rng)
        count += 1
        cur = _schedule(count, total, alpha, cur)
    return rng.pos


def train_bpr(pn, wv, wc, dim, iterations, total, alpha, lam, rng):
    """BPR.Train (bpr.go:61-131), one worker."""
    cur, count = alpha, 0
    for _ in range(iterations):
        user = pn.source_sample(rng)
        pos = pn.target_sample(user, rng)
        if pos == -1:
            continue
        neg = pn.negative_sample(rng)
        pn.update_bpr_pair(wv, wc, user, pos, neg, dim, cur, lam)
        count += 1
        cur = _schedule(count, total, alpha, cur)
    return rng.pos


def train_deepwalk(pn, wv, wc, dim, walk_times, walk_steps, window, K, alpha, rng, shuffle):
    """DeepWalk.Train (deepwalk.go:61-141), one worker; `shuffle` feeds the Fisher-Yates draws (rand.Int63n)."""
    total = walk_times * pn.max_vid
    cur, count, pairs = alpha, 0, 0
    for _ in range(walk_times):
        keys = list(range(pn.max_vid))
        for vid in range(pn.max_vid):
            j = vid + shuffle.intn(pn.max_vid - vid)
            keys[vid], keys[j] = keys[j], keys[vid]
        for vid in range(pn.max_vid):
            walk = pn.random_walk(keys[vid], walk_steps, rng)
            vs, cs = pn.skip_grams(walk, window)
            for v, c in zip(vs, cs):
                pn.update_pair(wv, wc, v, c, dim, K, cur, rng)
            pairs += len(vs)
            count += 1
            cur = _schedule(count, total, alpha, cur)
    return rng.pos, pairs


# ---- node2vec (internal/models/node2vec/node2vec.go) -----------------------------------------------------------------
def are_neighbors(pn, vid1, vid2):  # node2vec.go:165-173 (a linear scan there)
    return vid2 in pn.graph.get(vid1, [])


def biased_target_sample(pn, prev, current, p, q, rng):  # node2vec.go:113-162
    neighbors = pn.graph.get(current, [])
    if not neighbors:
        return -1
    weights = pn.edge[current]
    biased = []
    total = 0.0
    for i, nb in enumerate(neighbors):
        base = weights[i] if weights else 1.0
        if nb == prev:
            bias = 1.0 / p
        elif are_neighbors(pn, prev, nb):
            bias = 1.0
        else:
            bias = 1.0 / q
        biased.append(base * bias)
        total += biased[i]
    if total == 0:
        return neighbors[rng.intn(len(neighbors))]
    r = rng.float64() * total
    cum = 0.0
    for i, w in enumerate(biased):
        cum += w
        if r <= cum:
            return neighbors[i]
    return neighbors[-1]


def biased_random_walk(pn, start, steps, p, q, rng):  # node2vec.go:82-110
    walk = [start]
    if steps == 0:
        return walk
    first = pn.target_sample(start, rng)
    if first == -1:
        return walk
    walk.append(first)
    for _ in range(1, steps):
        nxt = biased_target_sample(pn, walk[-2], walk[-1], p, q, rng)
        if nxt == -1:
            break
        walk.append(nxt)
    return walk


def train_node2vec(pn, wv, wc, dim, walk_times, walk_steps, window, K, alpha, p, q, rng, shuffle):
    """Node2Vec.Train (node2vec.go:176-260), one worker: DeepWalk.Train with the biased walk."""
    total = walk_times * pn.max_vid
    cur, count, pairs = alpha, 0, 0
    for _ in range(walk_times):
        keys = list(range(pn.max_vid))
        for vid in range(pn.max_vid):
            j = vid + shuffle.intn(pn.max_vid - vid)
            keys[vid], keys[j] = keys[j], keys[vid]
        for vid in range(pn.max_vid):
            walk = biased_random_walk(pn, keys[vid], walk_steps, p, q, rng)
            vs, cs = pn.skip_grams(walk, window)
            for v, c in zip(vs, cs):
                pn.update_pair(wv, wc, v, c, dim, K, cur, rng)
            pairs += len(vs)
            count += 1
            cur = _schedule(count, total, alpha, cur)
    return rng.pos, pairs


# ---- CPR / TPR (Go tree only): a second graph and a third table --------------------------------------------------------
def cpr_transform_user(target, source, user_embed, target_embed, source_embed, user, dim):  # cpr.go:127-172
    out = [0.0] * dim
    count = 0.0
    for d in range(dim):
        out[d] += user_embed[user][d]
    count += 1.0
    for item in target.graph.get(user, []):  # userToTarget: vids of the target graph with neighbours (cpr.go:53-59)
        if item < len(target_embed):
            for d in range(dim):
                out[d] += target_embed[item][d]
            count += 1.0
    if user < source.max_vid:  # userToSource is filled for vid < sourceGraph.MaxVid only (cpr.go:72-78)
        for item in source.graph.get(user, []):
            if item < len(source_embed):
                for d in range(dim):
                    out[d] += source_embed[item][d]
                count += 1.0
    if count > 0:
        for d in range(dim):
            out[d] /= count
    return out


def train_cpr(target, source, user_embed, target_embed, source_embed, dim, iterations, total, alpha, user_reg, item_reg, margin, rng):
    """CPR.Train (cpr.go:175-282), one worker, `iterations` trips of a run of `total` samples. The source-domain item rows
    are only ever read. A trip whose user has no target-domain neighbour is skipped WITHOUT counting (cpr.go:214-216)."""
    cur, count = alpha, 0
    for _ in range(iterations):
        user = target.source_sample(rng)
        pos = target.target_sample(user, rng)
        if pos == -1:
            continue
        uv = cpr_transform_user(target, source, user_embed, target_embed, source_embed, user, dim)
        neg = target.negative_sample(rng)
        ps = ns = 0.0
        for d in range(dim):
            ps += uv[d] * target_embed[pos][d]
            ns += uv[d] * target_embed[neg][d]
        diff = ps - ns
        if diff < margin:
            g = cur * target.fast_sigmoid(-(diff - margin))
            pg, ng, ug = [0.0] * dim, [0.0] * dim, [0.0] * dim
            for d in range(dim):
                pg[d] = g * uv[d]
                ng[d] = -g * uv[d]
                ug[d] = g * (target_embed[pos][d] - target_embed[neg][d])
            for d in range(dim):
                user_embed[user][d] -= cur * user_reg * user_embed[user][d]
                user_embed[user][d] += ug[d]
                target_embed[pos][d] -= cur * item_reg * target_embed[pos][d]
                target_embed[pos][d] += pg[d]
                target_embed[neg][d] -= cur * item_reg * target_embed[neg][d]
                target_embed[neg][d] += ng[d]
        count += 1
        cur = _schedule(count, total, alpha, cur)
    return rng.pos


def tpr_item_vector(iw, item_embed, word_embed, item, dim, text_weight):  # tpr.go:101-121
    out = [(1.0 - text_weight) * item_embed[item][d] for d in range(dim)]
    words = iw.graph.get(item, [])
    if len(words) > 0:
        for word in words:
            for d in range(dim):
                out[d] += (text_weight / float(len(words))) * word_embed[word][d]
    else:
        out = [item_embed[item][d] for d in range(dim)]
    return out


def train_tpr(ui, iw, user_embed, item_embed, word_embed, dim, iterations, total, alpha, lam, text_weight, rng):
    """TPR.Train (tpr.go:124-262), one worker. Items are looked up in the item-word graph by their user-item vid."""
    cur, count = alpha, 0
    for _ in range(iterations):
        user = ui.source_sample(rng)
        pos = ui.target_sample(user, rng)
        if pos == -1:
            continue
        neg = ui.negative_sample(rng)
        pv = tpr_item_vector(iw, item_embed, word_embed, pos, dim, text_weight)
        nv = tpr_item_vector(iw, item_embed, word_embed, neg, dim, text_weight)
        ps = ns = 0.0
        for d in range(dim):
            ps += user_embed[user][d] * pv[d]
            ns += user_embed[user][d] * nv[d]
        coef = cur * ui.fast_sigmoid(ns - ps)
        ug, pg, ng = [0.0] * dim, [0.0] * dim, [0.0] * dim
        for d in range(dim):
            ug[d] = coef * (pv[d] - nv[d])
            pg[d] = coef * user_embed[user][d]
            ng[d] = -coef * user_embed[user][d]
        for d in range(dim):
            user_embed[user][d] += ug[d] - lam * cur * user_embed[user][d]
        for d in range(dim):
            item_embed[pos][d] += (1.0 - text_weight) * pg[d] - lam * cur * item_embed[pos][d]
            item_embed[neg][d] += (1.0 - text_weight) * ng[d] - lam * cur * item_embed[neg][d]
        pw = iw.graph.get(pos, [])
        if len(pw) > 0:
            ww = text_weight / float(len(pw))
            for word in pw:
                for d in range(dim):
                    word_embed[word][d] += ww * pg[d] - lam * cur * word_embed[word][d]
        nw = iw.graph.get(neg, [])
        if len(nw) > 0:
            ww = text_weight / float(len(nw))
            for word in nw:
                for d in range(dim):
                    word_embed[word][d] += ww * ng[d] - lam * cur * word_embed[word][d]
        count += 1
        cur = _schedule(count, total, alpha, cur)
    return rng.pos
