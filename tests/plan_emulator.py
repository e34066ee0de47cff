"""CPU interpreter of the PixelCNN execution plan exported by ts_debug_pixelcnn_plan.

TEST INFRASTRUCTURE.  Mirrors, in numpy, what csrc/pixelcnn.cu's executor does with the stage
table + packed weight blob (segment selection, ring buffers, epilogues), so the host-side packer
and the schedule can be validated against the oracle on a machine without a GPU.
"""
import numpy as np

(EPI_IDLE, EPI_VERT0, EPI_VERT, EPI_V2H, EPI_FUSEV, EPI_HGATE, EPI_HRES, EPI_FUSEH, EPI_OUT1, EPI_OUT2, EPI_SAMPLE,
 EPI_HRESF, EPI_HGATE2, EPI_OUT1F, EPI_V2H1) = range(15)
D, MB, SEG = 256, 64, 256 * 64


class Plan:
    def __init__(self, table, blob):
        h = table[:32]
        (self.ncta, self.nstages, self.L) = (int(h[0]), int(h[1]), int(h[2]))
        assert h[3] == D and h[4] == MB
        names = ["E", "XV1P", "XV", "HV", "V2H", "G", "XHP", "XH", "Y", "LOG", "CLS", "total"]
        self.lay = {n: int(h[5 + i]) for i, n in enumerate(names)}
        self.cl = max(1, int(h[18]))      # CTAs per work unit (cluster plan: the ranks hold K slices of the same rows)
        self.hvslots = int(h[19]) or 2    # pre-gate vertical outputs: 2-deep ring, or one slot per layer (schedule 2)
        self.table = table[32:].reshape(self.nstages, self.ncta, 8)
        self.blob = blob


def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x.astype(np.float32)))).astype(np.float32)


def hv_slot(layer, hvslots):
    return layer if hvslots > 2 else layer & 1


def segments(t, pas, r, lay, L, hvslots=2):
    epi, layer, col = int(t[0]), int(t[1]), int(t[2])
    if epi == EPI_V2H1:
        s0 = lay["HV"] + ((hv_slot(layer, hvslots) * 2 + col) * 2) * SEG
        return [s0, s0 + SEG]
    if epi == EPI_VERT0:
        return [lay["E"] + ((((r - 3 + kh) & 3) * 2) + ci) * SEG for kh in range(3) for ci in range(2)]
    if epi == EPI_VERT:
        return [lay["XV"] + ((layer * 2 + ((r - 1 + kh) & 1)) * 2 + ci) * SEG for kh in range(2) for ci in range(2)]
    if epi == EPI_V2H:
        s0 = lay["HV"] + ((hv_slot(layer, hvslots) * 2 + pas) * 2) * SEG
        return [s0, s0 + SEG]
    if epi == EPI_FUSEV:
        return [lay["XV1P"] + pas * SEG]
    if epi == EPI_HGATE:
        if layer == 0:
            return [] if col == 0 else [lay["E"] + ((r & 3) * 2 + 0) * SEG]
        s = [lay["XH"] + (0 * (L + 1) + layer) * SEG]
        if col == 1:
            s.append(lay["XH"] + (1 * (L + 1) + layer) * SEG)
        return s
    if epi == EPI_HRES:
        return [lay["G"] + (layer & 1) * SEG]
    if epi == EPI_HRESF:
        return [lay["G"]]
    if epi == EPI_HGATE2:
        s = [lay["G"] + ((layer - 1) & 1) * SEG, lay["XH"] + (col * (L + 1) + layer - 1) * SEG]
        if col == 1:
            s.append(lay["XH"] + (0 * (L + 1) + layer) * SEG)
        return s
    if epi == EPI_OUT1F:
        return [lay["G"] + ((L - 1) & 1) * SEG, lay["XH"] + (col * (L + 1) + L - 1) * SEG]
    if epi == EPI_FUSEH:
        return [lay["XHP"]]
    if epi == EPI_OUT1:
        return [lay["XH"] + (col * (L + 1) + L) * SEG]
    if epi == EPI_OUT2:
        return [lay["Y"], lay["Y"] + SEG]
    raise ValueError(epi)


def run(plan, emb, cls_w, audv, audh, label, codes_forced, T, noise=None, T0=None):
    """Interpret the plan for T rows.  Rows < T0 take codes_forced [B,T0,2]; later rows are sampled
    with noise [2(T-T0),B,2048] (argmax(softmax/noise)).  Returns (codes [B,T,2], logits [T,2,B,2048])."""
    B = len(label)
    T0 = T if T0 is None else T0
    lay, L = plan.lay, plan.L
    arena = np.zeros(lay["total"], dtype=np.float32)
    A = lambda off, n=SEG: arena[off:off + n].reshape(-1, MB)
    cls = arena[lay["CLS"]:lay["CLS"] + L * 2 * SEG].reshape(L, 2 * D, MB)
    for l in range(L):
        cls[l][:, :B] = cls_w[l][label].T
    codes = np.zeros((B, T, 2), dtype=np.int64)
    logits = np.zeros((T, 2, B, 2048), dtype=np.float32)
    for r in range(T):
        for s in range(plan.nstages):
            writes = []
            for cta in range(plan.ncta):
                t = plan.table[s, cta]
                epi, layer, col, row0, nrows, wofs, K, rpad = [int(v) for v in t]
                if epi == EPI_IDLE:
                    continue
                if epi == EPI_SAMPLE:
                    m = cta
                    if m >= B:
                        continue
                    lg = arena[lay["LOG"]:lay["LOG"] + 2048 * MB].reshape(2048, MB)[:, m].copy()
                    logits[r, col, m] = lg
                    if r < T0:
                        code = int(codes_forced[m, r, col])
                    else:
                        e = np.exp(lg - lg.max()).astype(np.float32)
                        p = (e / e.sum(dtype=np.float32)).astype(np.float32)
                        code = int(np.argmax(p / noise[2 * (r - T0) + col, m]))
                    codes[m, r, col] = code
                    writes.append((lay["E"] + ((r & 3) * 2 + col) * SEG, m, emb[code]))
                    if K:   # fused plan: layer-0 gate of column 1 gathered from the code table T0 [2048][512]
                        tw = plan.blob[wofs + code * 2 * D: wofs + (code + 1) * 2 * D]
                        v2h = A(lay["V2H"] + ((0 * 2 + 1) * 2) * SEG, 2 * SEG)
                        zt = (v2h[:D, m] + tw[0::2]) + cls[0][:D, m]
                        zs = (v2h[D:, m] + tw[1::2]) + cls[0][D:, m]
                        writes.append((lay["G"], m, np.tanh(zt) * _sigmoid(zs)))
                    continue
                cl = plan.cl
                if cta % cl:
                    continue                      # ranks 1.. of a cluster: their K slices are gathered by rank 0 below
                Ks = K // cl
                parts = []
                for q in range(cl):
                    tq = plan.table[s, cta + q]
                    assert [int(v) for v in tq[[0, 1, 2, 3, 4, 6, 7]]] == [epi, layer, col, row0, nrows, K, rpad]
                    wq = int(tq[5])
                    parts.append(plan.blob[wq:wq + Ks * rpad].reshape(Ks, rpad)[:, :nrows])
                    bq = plan.blob[wq + Ks * rpad: wq + Ks * rpad + nrows]
                    if q == 0:
                        bias = bq
                    else:
                        assert np.array_equal(bias, bq)
                W = np.concatenate(parts, 0) if K else np.zeros((0, nrows), np.float32)
                npass = 2 if epi in (EPI_V2H, EPI_FUSEV) else 1
                for pas in range(npass):
                    segs = segments(t, pas, r, lay, L, plan.hvslots)
                    assert len(segs) * D == K, (epi, layer, col, K, len(segs))
                    if K:
                        x = np.concatenate([A(o) for o in segs], 0)        # [K, MB]
                        acc = W.T.astype(np.float32) @ x
                    else:
                        acc = np.zeros((nrows, MB), np.float32)
                    acc = acc + bias[:, None]
                    if epi in (EPI_VERT0, EPI_VERT, EPI_HGATE, EPI_HGATE2):
                        q0, nq = row0 // 2, nrows // 2
                        at, as_ = acc[0::2], acc[1::2]
                        ct, cs = cls[layer][q0:q0 + nq], cls[layer][D + q0:D + q0 + nq]
                        if epi in (EPI_HGATE, EPI_HGATE2):
                            v2h = A(lay["V2H"] + ((layer * 2 + col) * 2) * SEG, 2 * SEG)
                            zt = (v2h[q0:q0 + nq] + at) + ct
                            zs = (v2h[D + q0:D + q0 + nq] + as_) + cs
                            writes.append((lay["G"] + (layer & 1) * SEG, (q0, nq), np.tanh(zt) * _sigmoid(zs)))
                        else:
                            hvo = lay["HV"] + ((hv_slot(layer, plan.hvslots) * 2 + col) * 2) * SEG
                            writes.append((hvo, (q0, nq), at.copy()))
                            writes.append((hvo, (D + q0, nq), as_.copy()))
                            g = np.tanh(at + ct) * _sigmoid(as_ + cs)
                            if epi == EPI_VERT0:
                                writes.append((lay["XV1P"] + col * SEG, (q0, nq), g))
                            elif layer + 1 < L:
                                writes.append((lay["XV"] + (((layer + 1) * 2 + (r & 1)) * 2 + col) * SEG, (q0, nq), g))
                    else:
                        sl = (row0, nrows)
                        if epi == EPI_V2H:
                            writes.append((lay["V2H"] + ((layer * 2 + pas) * 2) * SEG, sl, acc))
                        elif epi == EPI_V2H1:
                            writes.append((lay["V2H"] + ((layer * 2 + col) * 2) * SEG, sl, acc))
                        elif epi == EPI_FUSEV:
                            au = np.zeros((nrows, MB), np.float32)
                            au[:, :B] = audv[:, r, row0:row0 + nrows].T
                            writes.append((lay["XV"] + ((1 * 2 + (r & 1)) * 2 + pas) * SEG, sl, acc + au))
                        elif epi == EPI_HRES:
                            if layer == 0:
                                writes.append((lay["XHP"], sl, acc))
                            else:
                                xh = A(lay["XH"] + (col * (L + 1) + layer) * SEG)[row0:row0 + nrows]
                                writes.append((lay["XH"] + (col * (L + 1) + layer + 1) * SEG, sl, acc + xh))
                        elif epi in (EPI_FUSEH, EPI_HRESF):
                            au = np.zeros((nrows, MB), np.float32)
                            au[:, :B] = audh[:, r, row0:row0 + nrows].T
                            writes.append((lay["XH"] + (col * (L + 1) + 1) * SEG, sl, acc + au))
                        elif epi in (EPI_OUT1, EPI_OUT1F):
                            writes.append((lay["Y"], sl, np.maximum(acc, 0)))
                        elif epi == EPI_OUT2:
                            writes.append((lay["LOG"], sl, acc))
            # stage barrier: apply all writes after every task of the stage has read its inputs
            for off, where, val in writes:
                if isinstance(where, tuple):
                    r0, n = where
                    arena[off + r0 * MB: off + (r0 + n) * MB] = val.astype(np.float32).reshape(-1)
                else:
                    arena[off + where: off + SEG: MB] = val            # column m of a [256][MB] segment
    return codes, logits
