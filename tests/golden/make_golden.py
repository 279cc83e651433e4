"""Writes tests/golden/kat.json: known-answer vectors for the hot path.

The reference ships no golden vectors (PARITY UNPINNED), so these are derived here
WITHOUT the oracle: quantization cases by hand from compute/quantization.go (the arithmetic is written
out in each `why`), dequantization / cosine cases by plain Python float arithmetic (IEEE-754 binary64,
round-to-nearest, no fusion) following compute/cosine.go line by line.  Both oracles (oracle.c and
oracle_np.py) and the CUDA path must reproduce them bit-for-bit.

Run: python tests/golden/make_golden.py   (rewrites kat.json deterministically)
"""
import json
import math
import os
import struct


def f32_le(x):
    return list(struct.pack("<f", x))


def row(mn, mx, codes):
    return f32_le(mn) + f32_le(mx) + list(codes)


def f64_hex(x):
    return struct.pack(">d", x).hex()


def f32_hex(x):
    return struct.pack(">f", x).hex()


quantize_f32 = [
    dict(name="positive_keeps_min_zero", input=[0.5, 1.0, 0.25], expect=row(0.0, 1.0, [127, 255, 63]),
         why="range seeded at 0 -> min=0,max=1; 0.5*255=127.5 -> 127 (truncation), 1.0 -> 255, 0.25*255=63.75 -> 63"),
    dict(name="symmetric", input=[-1.0, 0.0, 1.0], expect=row(-1.0, 1.0, [0, 127, 255]),
         why="(v+1)/2*255: 0, 127.5 -> 127, 255"),
    dict(name="all_positive", input=[2.0, 4.0], expect=row(0.0, 4.0, [127, 255]),
         why="min stays 0 (seed), 2/4*255=127.5 -> 127"),
    dict(name="all_negative", input=[-2.0, -4.0], expect=row(-4.0, 0.0, [127, 0]),
         why="max stays 0 (seed), (-2+4)/4*255=127.5 -> 127, -4 -> 0"),
    dict(name="all_zero", input=[0.0, 0.0, 0.0], expect=row(0.0, 0.0, [0, 0, 0]),
         why="0/0 = NaN; uint8(NaN*255) is 0 on amd64 (CVTTSS2SL low byte)"),
    dict(name="nan_ignored_by_range", input=[float("nan"), 1.0], expect=row(0.0, 1.0, [0, 255]),
         why="NaN fails both comparisons in rangeFloat32 and both clamps; NaN code -> 0"),
    dict(name="truncation_not_rounding", input=[0.0, 1.0, 0.999], expect=row(0.0, 1.0, [0, 255, 254]),
         why="0.999*255 = 254.745 -> 254"),
    dict(name="single_positive", input=[5.0], expect=row(0.0, 5.0, [255]), why="5/5*255"),
    dict(name="single_negative", input=[-5.0], expect=row(-5.0, 0.0, [0]), why="(-5+5)/5*255 = 0"),
    dict(name="empty", input=[], expect=row(0.0, 0.0, []), why="8-byte header of zeros, no codes"),
    dict(name="value_equals_max_is_255", input=[-0.25, 0.75, 0.75], expect=row(-0.25, 0.75, [0, 255, 255]),
         why="(0.75+0.25)/1.0*255 = 255 exactly"),
]

quantize_f64 = [
    dict(name="header_is_float32_rounded_codes_use_float64_range", input=[0.1, -0.3],
         expect=row(0.1, -0.3, [])[4:8] + row(0.1, -0.3, [])[0:4] + [255, 0],
         why="header = float32(-0.3), float32(0.1); (0.1-(-0.3))/(0.1-(-0.3)) = 1.0 in float64 -> 255; min -> 0"),
    dict(name="f64_truncation", input=[0.0, 1.0, 0.999], expect=row(0.0, 1.0, [0, 255, 254]), why="as float32 case"),
    dict(name="f64_all_zero", input=[0.0, 0.0], expect=row(0.0, 0.0, [0, 0]), why="NaN -> 0 (CVTTSD2SQ low byte)"),
]
# the first f64 case above builds the header by hand: float32(-0.3) then float32(0.1)
quantize_f64[0]["expect"] = f32_le(-0.3) + f32_le(0.1) + [255, 0]


def deq64(q, mn, mx):
    normalized = float(q) / 255.0
    return mn + normalized * (mx - mn)


def deq32(q, mn, mx):
    f = lambda x: struct.unpack("<f", struct.pack("<f", x))[0]  # round to float32
    normalized = f(float(q) / 255.0)   # float32 division of exactly representable operands
    rng = f(mx - mn)
    return f(mn + f(normalized * rng))


dequantize = []
for (mn, mx, codes) in [(-1.0, 1.0, [0, 255, 51, 127, 128]), (0.0, 1.0, [0, 1, 254, 255]), (-0.125, 0.375, [17, 200])]:
    dequantize.append(dict(row=row(mn, mx, codes),
                           f64=[f64_hex(deq64(q, mn, mx)) for q in codes],
                           f32=[f32_hex(deq32(q, mn, mx)) for q in codes]))


def normalize(v):
    norm = 0.0
    for x in v:
        norm += x * x
    norm = math.sqrt(norm)
    if norm != 0:
        v = [x / norm for x in v]
    return v


def cosine(qrow, r):
    d = len(qrow) - 8
    qmn, qmx = struct.unpack("<ff", bytes(qrow[:8]))
    rmn, rmx = struct.unpack("<ff", bytes(r[:8]))
    A = normalize([deq64(qrow[8 + i], qmn, qmx) for i in range(d)])
    B = normalize([deq64(r[8 + i], rmn, rmx) for i in range(d)])
    dot = 0.0
    for a, b in zip(A, B):
        dot += a * b
    return struct.unpack("<f", struct.pack("<f", dot))[0]


q = row(0.0, 1.0, [255, 0])
cos_rows = [row(0.0, 1.0, [0, 255]), row(0.0, 1.0, [255, 255]), row(0.0, 1.0, [255, 0]), row(0.0, 0.0, [7, 9]),
            row(-1.0, 1.0, [0, 255]), row(-1.0, 0.0, [0, 255])]
cosine_cases = dict(
    query=q, rows=cos_rows,
    dots=[0, 255 * 255, 255 * 255, 255 * 7, 0, 0],
    sims_f32=[f32_hex(cosine(q, r)) for r in cos_rows],
    why=["orthogonal -> 0", "(1,1)/sqrt2 . (1,0) = 0.70710677", "identical -> 1",
         "header 0/0: every value dequantizes to 0, norm 0, left unnormalized, dot = +0",
         "(-1,1)/sqrt2 . (1,0) = -0.70710677", "(-1,0) . (1,0) = -1"])

argmax_cases = dict(
    centroids=[row(0.0, 1.0, [255, 0]), row(0.0, 1.0, [0, 255]), row(0.0, 1.0, [255, 0])],
    data=[row(0.0, 1.0, [255, 0]), row(0.0, 1.0, [0, 255]), row(0.0, 0.0, [0, 0]), row(0.0, 1.0, [255, 255]),
          row(-1.0, 0.0, [0, 255])],
    argmax=[0, 1, 0, 0, 1],
    why=["tie between identical centroids 0 and 2: strict '>' keeps the lowest index", "centroid 1",
         "zero row: every dot is +0 > -1.0 -> index 0", "tie 0.7071 between 0 and 1 -> lowest index",
         "(-1,0): dots are -1,0,-1 -> centroid 1"])

# server/search.go:202-273 and server/upload.go:239-279 on five rows behind two centroids, by hand.
# Centroid of each row (cosine.go:70-125): (0,1) -> 1; (1,1) ties 0.7071 -> lowest index 0; (1,0) -> 0; the zero row scores
# +0 > -1.0 against centroid 0 -> 0; (-1,0) scores -1 and 0 -> 1.  Query (1,0): centroid sims 1 and 0.
s_cent = [row(0.0, 1.0, [255, 0]), row(0.0, 1.0, [0, 255])]
s_rows = [row(0.0, 1.0, [0, 255]), row(0.0, 1.0, [255, 255]), row(0.0, 1.0, [255, 0]), row(0.0, 0.0, [0, 0]),
          row(-1.0, 0.0, [0, 255])]
s_doc = [100, 101, 102, 103, 104]
s_lists = [1, 0, 0, 0, 1]
s_sim = {d: f32_hex(cosine(q, r)) for d, r in zip(s_doc, s_rows)}
search_cases = dict(
    centroids=s_cent, rows=s_rows, doc_ids=s_doc, lists=s_lists, query=q,
    searches=[
        dict(nprobe=1, k=10, ids=[102, 101, 103], sims_f32=[s_sim[102], s_sim[101], s_sim[103]],
             why="only list 0 is probed (centroid sims 1 > 0): its rows score 0.7071, 1, +0 -> sorted by similarity desc"),
        dict(nprobe=2, k=10, ids=[102, 101, 100, 103, 104], sims_f32=[s_sim[i] for i in (102, 101, 100, 103, 104)],
             why="every list: 1, 0.7071, then documents 100 and 103 tie at +0 -> lower id first, then -1"),
        dict(nprobe=2, k=2, ids=[102, 101], sims_f32=[s_sim[102], s_sim[101]], why="truncated to Count+Offset = 2"),
    ],
    upload=dict(first=3, assign=[0, 1],
                why="the index holds rows 0-2; rows 3 and 4 are uploaded: the zero row goes to centroid 0, (-1,0) to centroid 1 "
                    "(upload.go:245); the index then answers like the one built from all five rows"))
same_doc = dict(doc_ids=[7, 8, 7, 9, 8], nprobe=2, k=10, ids=[7, 8, 9], sims_f32=[s_sim[102], s_sim[101], s_sim[103]],
                why="search.go:260-268: one hit per document keeping its best embedding: document 7 (rows 0, 2) keeps 1, "
                    "document 8 (rows 1, 4) keeps 0.7071, document 9 has +0")
search_cases["dedup"] = same_doc



# ---- dnc/k_means.go:67-117 (one Lloyd iteration) and dnc/dnc.go:417-449 (recenter), in plain Python arithmetic ----
def f32(x):
    return struct.unpack("<f", struct.pack("<f", x))[0]


def quantize32(vec):
    """quantization.go:21-45,82-91,194-216 in float32: range seeded at 0, clamp, ((v-min)/(max-min))*255 truncated."""
    mn = mx = 0.0
    for v in vec:
        if v < mn:
            mn = v
        if v > mx:
            mx = v
    codes = []
    for v in vec:
        v = min(max(v, mn), mx)
        if mx == mn:
            codes.append(0)          # 0/0 = NaN -> uint8(NaN) = 0 on amd64
            continue
        normalized = f32(f32(v - mn) / f32(mx - mn))
        codes.append(int(f32(normalized * 255.0)) & 0xFF)
    return row(mn, mx, codes)


def quantize64(vec):
    """quantization.go:8-19,93-102 in float64; the header stores float32(min), float32(max)."""
    mn = mx = 0.0
    for v in vec:
        if v < mn:
            mn = v
        if v > mx:
            mx = v
    codes = []
    for v in vec:
        v = min(max(v, mn), mx)
        codes.append(0 if mx == mn else int(((v - mn) / (mx - mn)) * 255.0) & 0xFF)
    return row(mn, mx, codes)


def dequantize_row(r, fn):
    mn, mx = struct.unpack("<ff", bytes(r[:8]))
    return [fn(c, mn, mx) for c in r[8:]]


km_cent = [row(0.0, 1.0, [255, 0]), row(0.0, 1.0, [0, 255]), row(-1.0, 0.0, [0, 255])]
km_data = [row(0.0, 1.0, [255, 0]), row(0.0, 1.0, [200, 100]), row(0.0, 1.0, [0, 255]), row(0.0, 1.0, [100, 100]),
           row(0.0, 0.0, [9, 9])]
km_assign = [0, 0, 1, 0, 0]   # (1,0); cos .894 vs .447; (0,1); tie .7071 -> lowest index; zero row: +0 > -1.0 -> index 0
km_prev_means = [[0.5, 0.5], [0.125, 0.0], [0.25, 0.5]]
km_sums = [[0.0, 0.0] for _ in km_cent]
km_counts = [0, 0, 0]
for i, c in enumerate(km_assign):                       # k_means.go:80-86: float32 sums in row order
    vec = dequantize_row(km_data[i], deq32)
    for j, val in enumerate(vec):
        km_sums[c][j] = f32(km_sums[c][j] + val)
    km_counts[c] += 1
km_means = [list(m) for m in km_prev_means]
for c in range(len(km_cent)):                            # k_means.go:89-96: an empty cluster keeps its previous mean
    if km_counts[c] > 0:
        km_means[c] = [f32(sv / f32(float(km_counts[c]))) for sv in km_sums[c]]
km_new = [quantize32(m) for m in km_means]              # k_means.go:99
kmeans_case = dict(
    centroids=km_cent, data=km_data, prev_means=km_prev_means, assign=km_assign, counts=km_counts,
    means_f32=[[f32_hex(v) for v in m] for m in km_means], new_centroids=km_new,
    converged=all(n[8:] == c[8:] for n, c in zip(km_new, km_cent)),
    why="cluster 0 = rows 0,1,3,4 (the zero row adds +0 but counts), cluster 1 = row 2, cluster 2 empty: it keeps the mean it "
        "had, (0.25, 0.5), and is requantized from it; sums and means in float32, row order")

rc_rows = [row(0.0, 1.0, [255, 0]), row(0.0, 1.0, [0, 255]), row(-1.0, 0.0, [0, 255])]
rc_sum = [0.0, 0.0]
for r in rc_rows:                                        # dnc.go:417-436: float64 sums in row order
    for j, val in enumerate(dequantize_row(r, deq64)):
        rc_sum[j] += val
rc_mean = [v / float(len(rc_rows)) for v in rc_sum]      # dnc.go:446-448
recenter_case = dict(rows=rc_rows, expect=quantize64(rc_mean),
                     why="(1,0) + (0,1) + (-1,0) = (0,1); / 3 = (0, 0.333..): range [0, 1/3] (header float32(1/3)), codes 0 and 255")

out = dict(quantize_f32=quantize_f32, quantize_f64=quantize_f64, dequantize=dequantize, cosine=cosine_cases,
           argmax=argmax_cases, search=search_cases, kmeans_step=kmeans_case, recenter=recenter_case)


def clean(o):
    if isinstance(o, float) and o != o:
        return "nan"
    if isinstance(o, dict):
        return {k: clean(v) for k, v in o.items()}
    if isinstance(o, list):
        return [clean(v) for v in o]
    return o


with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat.json"), "w") as f:
    json.dump(clean(out), f, indent=1)
print("wrote kat.json")
