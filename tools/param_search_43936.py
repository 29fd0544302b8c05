"""Search for a structure of the missing GAT3D model with the one upstream known answer:
GATMultistream.Model(image_width=20, image_height=20, n_vertices=6, attention_type="temporal", mapping_type="conv")
has 43,936 trainable parameters (reference convolutional_gat/compare_models/results/results.json:9).

Building blocks follow what the in-tree siblings and call sites make plausible (baseline_model.py:13-25,105-117: W, a, B
per head; model.py:21-42: layers get nfeat = nhid = time_steps = 4, image_height, image_width, n_vertices; heads 3 + 1):
  per head:   a projection (Conv2d ci->co, k x k, bias or not; or Linear), an attention vector `a`, an adjacency `B`,
              each optionally per pixel (image_height x image_width are constructor arguments of the layer) or per vertex;
  per model:  `heads_total` heads (hidden + output layers, times streams), optional extra convs.
Prints every combination that hits 43,936 exactly.  Pure Python, seconds.
"""
import itertools

H = W = 20
V, T = 6, 4
P = H * W
TARGET = 43936


def conv(ci, co, k, bias):
    return co * ci * k * k + (co if bias else 0)


projections = {}
for ci, co in [(4, 4), (6, 6), (1, 1), (24, 24), (4, 12), (6, 18), (24, 72)]:
    for k in (1, 3, 5, 7):
        for bias in (True, False):
            projections[f"Conv2d({ci}->{co},k{k}{',bias' if bias else ''})"] = conv(ci, co, k, bias)
for ci, co, k in [(4, 4, 3), (6, 6, 3)]:
    projections[f"{V}x per-vertex Conv2d({ci}->{co},k{k},bias)"] = V * conv(ci, co, k, True)
    projections[f"{T}x per-frame Conv2d({ci}->{co},k{k},bias)"] = T * conv(ci, co, k, True)
projections["Conv3d(1->1,k3,bias)"] = 28
projections["Conv3d(4->4,k3,bias)"] = 4 * 4 * 27 + 4
projections["Conv3d(6->6,k3,bias)"] = 6 * 6 * 27 + 6
projections["Linear(4->4) W"] = 16
projections["Linear(6->6) W"] = 36

a_sizes = {"a[2*4]": 8, "a[2*6]": 12, "a[2*4] per pixel": 8 * P, "a[2*6] per pixel": 12 * P, "a[2*4*V]": 48, "a[2*6*T]": 48,
           "a[2*P*4] (1-D layer on flattened maps)": 2 * P * 4, "a[2*P*6]": 2 * P * 6, "a[2*P]": 2 * P, "a[2]": 2,
           "a[2*4] per vertex per pixel": 8 * P * V, "none": 0}
b_sizes = {"B[V,V]": V * V, "B[T,T]": T * T, "B[V,V] per pixel": V * V * P, "B[T,T] per pixel": T * T * P,
           "B[V,V] per frame": V * V * T, "B[T,T] per vertex": T * T * V, "B[P,P]": P * P, "B[V,V] per 2x2 block": V * V * P // 4,
           "B[T,T] per 2x2 block": T * T * P // 4, "B[V,V]+B[T,T]": V * V + T * T, "none": 0}
extras = {"none": 0, "out Conv2d(4->4,k3,bias)": 148, "out Conv2d(4->4,k1,bias)": 20, "out Conv2d(24->24,k1,bias)": 600,
          "out Conv2d(12->4,k1,bias)": 52, "out Conv2d(12->4,k3,bias)": 436, "out Linear(12->4)+bias": 52,
          "BatchNorm2d(4)": 8, "BatchNorm2d(24)": 48, "per-pixel bias [P,T,V]": P * T * V, "per-pixel scale+bias": 2 * P * T * V,
          "stream weights [3]": 3, "stream weights [2]": 2}

hits = []
for heads in (1, 2, 3, 4, 6, 7, 8, 9, 12, 16):
    for (pn, pv), (an, av), (bn, bv), (en, ev) in itertools.product(projections.items(), a_sizes.items(), b_sizes.items(),
                                                                    extras.items()):
        if heads * (pv + av + bv) + ev == TARGET:
            hits.append((heads, pn, an, bn, en))
print(f"{len(hits)} structure(s) with exactly {TARGET} parameters at {H}x{W}, V={V}, T={T}:")
for h in hits:
    print("  heads=%d  projection=%s  attention=%s  adjacency=%s  extra=%s" % h)
# the structure built in this repo (cgat.layers / convolutional_gat.GAT3D), for the record
ours = 4 * (conv(6, 6, 3, True) + 12 + 16)
print(f"this repo's temporal/conv model (3 + 1 heads of Conv2d(6->6,k3,bias) + a[12] + B[4,4]): {ours}")
