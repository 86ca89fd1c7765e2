"""A synthetic AirV2X directory tree (`<root>/<scenario>/timestamp_%06d/{objects.pkl, agent_%06d/...}`) in the layout
`parse_seq` (utils/airv2x_utils.py:121-263) reads: metadata / object pickles, binary .pcd clouds with the intensity packed
into the rgb field (as open3d writes them), camera / depth / map PNGs. TEST INFRASTRUCTURE."""
import os
import pickle

import numpy as np

CAMS = {"vehicle": ["front", "front_left", "front_right", "rear", "rear_left", "rear_right"],
        "rsu": ["back", "front", "left", "right"], "drone": ["bev"]}


def write_pcd(path, cloud):
    """x y z rgb, DATA binary; rgb = the float whose bits are (r << 16 | g << 8 | b), r = round(intensity * 255)"""
    n = cloud.shape[0]
    r = np.clip(np.round(cloud[:, 3] * 255), 0, 255).astype(np.uint32)
    rec = np.zeros(n, dtype=[("x", "f4"), ("y", "f4"), ("z", "f4"), ("rgb", "f4")])
    rec["x"], rec["y"], rec["z"] = cloud[:, 0], cloud[:, 1], cloud[:, 2]
    rec["rgb"] = (r << 16).view(np.float32)
    head = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F F\n"
            "COUNT 1 1 1 1\nWIDTH %d\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary\n" % (n, n))
    with open(path, "wb") as f:
        f.write(head.encode("ascii"))
        f.write(rec.tobytes())


def write_tree(root, seed=0, n_scenarios=2, n_timestamps=3, n_pts=600, img_hw=(36, 64), late_agent=True):
    """agents: ids are NOT grouped by type (an RSU has the smallest id); one vehicle only exists from the 2nd timestamp on"""
    from PIL import Image
    rng = np.random.default_rng(seed)
    layout = [(3, "rsu"), (11, "vehicle"), (12, "drone"), (20, "vehicle"), (21, "rsu"), (35, "vehicle")]
    for s in range(n_scenarios):
        sc = os.path.join(root, "2025_01_0%d_00_00_00" % (s + 1))
        base = {a: [rng.uniform(-40, 40), rng.uniform(-40, 40), {"vehicle": 0.3, "rsu": 4.0, "drone": 40.0}[k], 0.0,
                    rng.uniform(-180, 180), 0.0] for a, k in layout}
        for t in range(n_timestamps):
            ts = os.path.join(sc, "timestamp_%06d" % (t * 5))
            os.makedirs(ts)
            objects = {}
            for k in range(30):
                objects[500 + k] = {"location": [rng.uniform(-120, 120), rng.uniform(-80, 80), rng.uniform(-0.2, 0.2),
                                                 0.0, rng.uniform(-180, 180), 0.0],
                                    "center": [0.0, 0.0, rng.uniform(0.6, 0.9)],
                                    "extent": [rng.uniform(1.6, 2.6), rng.uniform(0.7, 1.0), rng.uniform(0.7, 0.9)],
                                    "class": int(rng.integers(0, 9))}          # 0, 7, 8 are filtered out
            with open(os.path.join(ts, "objects.pkl"), "wb") as f:
                pickle.dump(objects, f)
            for a, kind in layout:
                if late_agent and a == 35 and t == 0 and s == 0:
                    continue
                d = os.path.join(ts, "agent_%06d" % a)
                os.makedirs(d)
                pos = list(base[a])
                if kind != "rsu":
                    pos[0] += 1.2 * t
                    pos[4] += 2.0 * t
                meta = {"agent_type": kind, "lidar": {"lidar_pose": [0.0, 0.0, 1.9 if kind != "drone" else -0.5, 0.0, 0.0, 0.0]},
                        "odometry": {"ego_pos": pos, "ego_speed": float(rng.uniform(0, 20))}}
                for c in CAMS[kind]:
                    ext = np.eye(4)
                    yaw = np.radians(rng.uniform(-180, 180))
                    ext[:2, :2] = [[np.cos(yaw), -np.sin(yaw)], [np.sin(yaw), np.cos(yaw)]]
                    ext[:3, 3] = rng.uniform(-1, 1, 3)
                    meta[c + "_camera"] = {"intrinsic": [[640.0, 0, 640.0], [0, 640.0, 360.0], [0, 0, 1.0]],
                                           "extrinsic": ext.tolist(), "cords": [0.5, 0.0, 1.5, 0.0, 0.0, 0.0]}
                    Image.fromarray(rng.integers(0, 256, img_hw + (3,), dtype=np.uint8)).save(os.path.join(d, c + "_camera.png"))
                    Image.fromarray(rng.integers(0, 256, img_hw + (3,), dtype=np.uint8)).save(os.path.join(d, c + "_depth.png"))
                with open(os.path.join(d, "metadata.pkl"), "wb") as f:
                    pickle.dump(meta, f)
                cloud = np.stack([rng.normal(0, 35, n_pts), rng.normal(0, 15, n_pts), rng.uniform(-3, 1, n_pts) - (pos[2] - 0.3),
                                  rng.integers(0, 256, n_pts) / 255.0], axis=1).astype(np.float32)
                write_pcd(os.path.join(d, "lidar.pcd"), cloud)
                for name in ["map_static_background.png", "map_static_lane.png", "map_static_road.png"] + \
                        ["map_dynamic_bev_layer_%d.png" % i for i in range(7)]:
                    Image.fromarray((rng.uniform(0, 1, (16, 24)) > 0.7).astype(np.uint8) * 255).save(os.path.join(d, name))
                with open(os.path.join(d, "vector_map.json"), "w") as f:
                    f.write("{}")
    return root


def lzf_compress(data):
    """a small greedy LZF encoder (longest match in the last 8191 bytes, O(n * window): test-sized inputs only)"""
    data = bytes(data)
    out, lit, i, n = bytearray(), bytearray(), 0, len(data)

    def flush():
        for k in range(0, len(lit), 32):
            chunk = lit[k:k + 32]
            out.append(len(chunk) - 1)
            out.extend(chunk)
        lit.clear()
    while i < n:
        best, dist = 0, 0
        if i + 3 <= n:
            start = max(0, i - 8191)
            j = data.rfind(data[i:i + 3], start, i + 2)
            while j != -1 and j < i:
                m = 3
                while i + m < n and m < 264 and data[j + m] == data[i + m]:
                    m += 1
                if m > best:
                    best, dist = m, i - j
                j = data.rfind(data[i:i + 3], start, j + 2) if j > start else -1
        if best >= 3:
            flush()
            length, off = best - 2, dist - 1
            if length < 7:
                out.append((length << 5) | (off >> 8))
            else:
                out.append((7 << 5) | (off >> 8))
                out.append(length - 7)
            out.append(off & 255)
            i += best
        else:
            lit.append(data[i])
            i += 1
    flush()
    return bytes(out)


def write_pcd_compressed(path, cloud):
    """DATA binary_compressed: the fields one after the other, LZF-compressed, preceded by the two sizes"""
    n = cloud.shape[0]
    r = np.clip(np.round(cloud[:, 3] * 255), 0, 255).astype(np.uint32)
    flat = b"".join([cloud[:, 0].astype("<f4").tobytes(), cloud[:, 1].astype("<f4").tobytes(),
                     cloud[:, 2].astype("<f4").tobytes(), (r << 16).astype("<u4").tobytes()])
    comp = lzf_compress(flat)
    head = ("VERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F U\nCOUNT 1 1 1 1\nWIDTH %d\nHEIGHT 1\n"
            "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS %d\nDATA binary_compressed\n" % (n, n))
    with open(path, "wb") as f:
        f.write(head.encode("ascii"))
        f.write(np.array([len(comp), len(flat)], dtype="<u4").tobytes())
        f.write(comp)
