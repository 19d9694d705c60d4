"""Generates tests/golden/plate_c1_input.npz (INPUT data of BASELINE config C1).

Run in the build container only (reads /root/reference, which does not exist on
the GPU box):   python tests/golden/make_plate_inputs.py

Sources (input data, not results):
  /root/reference/demos_csdl_alpha/thickness_opt/geometry/plate_geometry.igs
      six IGES type-128 bicubic B-spline surfaces (unit weights)
  /root/reference/demos_csdl_alpha/thickness_opt/plate_int_data.npz
      intersection cache written by PENGoLINS ``save_intersections_data``
      (name1 n_int, name2 mapping_list, name3 physical coords,
       name4 parametric coords per side, name5 curve length, name6 mortar_nels)
"""
import re
import numpy as np

SRC = "/root/reference/demos_csdl_alpha/thickness_opt/"


def parse_iges_128(path):
    pdata = {}
    for ln in open(path):
        if len(ln) >= 73 and ln[72] == "P":
            de = int(ln[64:72])
            pdata.setdefault(de, []).append(ln[:64])
    surfs = []
    for de in sorted(pdata):
        txt = "".join(pdata[de]).replace(" ", "").rstrip(";")
        txt = txt.split(";")[0]
        f = txt.split(",")
        if f[0] != "128":
            continue
        v = [float(x.replace("D", "E")) if x else 0.0 for x in f[1:]]
        K1, K2, M1, M2 = int(v[0]), int(v[1]), int(v[2]), int(v[3])
        o = 9
        n1, n2 = K1 + M1 + 2, K2 + M2 + 2
        ku = np.array(v[o:o + n1]); o += n1
        kv = np.array(v[o:o + n2]); o += n2
        nw = (K1 + 1) * (K2 + 1)
        w = np.array(v[o:o + nw]); o += nw
        cp = np.array(v[o:o + 3 * nw]).reshape(nw, 3); o += 3 * nw  # u fastest
        surfs.append(dict(p=(M1, M2), ku=ku, kv=kv, w=w, cp=cp, shape=(K1 + 1, K2 + 1)))
    return surfs


if __name__ == "__main__":
    surfs = parse_iges_128(SRC + "geometry/plate_geometry.igs")
    d = np.load(SRC + "plate_int_data.npz", allow_pickle=True)
    out = {"num_patches": len(surfs)}
    for s, S in enumerate(surfs):
        out[f"p{s}_deg"] = np.array(S["p"])
        out[f"p{s}_ku"] = S["ku"]; out[f"p{s}_kv"] = S["kv"]
        out[f"p{s}_cp"] = np.concatenate([S["cp"] * S["w"][:, None], S["w"][:, None]], axis=1)
        print(s, S["p"], S["shape"], S["cp"].min(0), S["cp"].max(0), S["w"].min(), S["w"].max())
    out["mapping_list"] = np.asarray(d["name2"], dtype=np.int64)
    out["mortar_nels"] = np.asarray(d["name6"], dtype=np.int64)
    for i in range(int(d["name1"])):
        for side in range(2):
            out[f"int{i}_xi{side}"] = np.asarray(d["name4"][i][side], dtype=np.float64)
        out[f"int{i}_phys"] = np.asarray(d["name3"][i], dtype=np.float64)
        print(i, out[f"int{i}_xi0"].shape, out[f"int{i}_xi0"][[0, -1]], out[f"int{i}_xi1"][[0, -1]])
    np.savez_compressed(__file__.replace("make_plate_inputs.py", "plate_c1_input.npz"), **out)
