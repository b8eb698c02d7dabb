"""Fabricates a tiny CATER-format data set in the reference's own file formats:
    <root>/od_perception/<video>.pkl      {"bb": [int[n,4] per frame], "labels": [int[n] per frame]}   (detector output)
    <root>/labels/<video>_bb.json         {"small_gold_spl_metal_Spl_0": [[x, y, w, h] x 300], ...}     (ground truth)
    <root>/containment_annotations.txt    "<video>\\tf1,f2,..."                                          (mask frames)
    <root>/videos/<video>.avi             301 black 320x240 frames (only when with_videos: the reference's inference
                                          writes its _bb.json files from inside the cv2 video loop)
and the matching training / inference config dictionaries of configs/training_config.json / inference_config.json."""
import json
import os
import pickle

import numpy as np

SNITCH, SNITCH_NAME = 140, "small_gold_spl_metal_Spl_0"
CONES, OTHERS = [0, 4, 8, 12], [1, 2, 3, 5, 6, 7, 9, 10]


def fabricate(root: str, n_videos: int = 3, seed: int = 0, with_videos: bool = False, T: int = 300):
    rng = np.random.RandomState(seed)
    for sub in ("od_perception", "labels", "videos"):
        os.makedirs(os.path.join(root, sub), exist_ok=True)
    names = [f"CATER_new_{i:06d}" for i in range(n_videos)]
    lines = []
    for name in names:
        objects = [SNITCH] + sorted(rng.choice(CONES, 2, replace=False).tolist() + rng.choice(OTHERS, 4, replace=False).tolist())
        start = {o: rng.randint(20, 200, size=2) for o in objects}
        vel = {o: rng.uniform(-0.3, 0.3, size=2) for o in objects}
        hidden = sorted(rng.choice(np.arange(40, T - 40), 60, replace=False).tolist())   # snitch not detected there
        bbs, labs, snitch_track = [], [], []
        for t in range(T):
            frame_bb, frame_lab = [], []
            for o in objects:
                x, y = (start[o] + vel[o] * t).astype(int)
                w, h = (12, 14) if o == SNITCH else (30, 34)
                box = [int(x), int(y), int(x + w), int(y + h)]
                if o == SNITCH:
                    snitch_track.append([int(x), int(y), w, h])
                    if t in hidden:
                        continue
                elif rng.rand() < 0.1:
                    continue                                                             # a missed detection
                frame_bb.append(box)
                frame_lab.append(o)
            bbs.append(np.array(frame_bb, dtype=np.int64).reshape(-1, 4))
            labs.append(np.array(frame_lab, dtype=np.int64))
        with open(os.path.join(root, "od_perception", name + ".pkl"), "wb") as f:
            pickle.dump({"bb": bbs, "labels": labs}, f)
        with open(os.path.join(root, "labels", name + "_bb.json"), "w") as f:
            json.dump({SNITCH_NAME: snitch_track}, f)
        lines.append(name + "\t" + ",".join(str(t) for t in hidden) + "\n")
        if with_videos:
            import cv2
            writer = cv2.VideoWriter(os.path.join(root, "videos", name + ".avi"), cv2.VideoWriter_fourcc(*"MJPG"), 24, (320, 240))
            frame = np.zeros((240, 320, 3), np.uint8)
            for _ in range(T + 1):
                writer.write(frame)
            writer.release()
    with open(os.path.join(root, "containment_annotations.txt"), "w") as f:
        f.writelines(lines)
    return names


def training_config(root: str, device: str, checkpoints: str, batch_size: int = 2, epochs: int = 1):
    data = {"sample_dir": os.path.join(root, "od_perception"), "labels_dir": os.path.join(root, "labels"),
            "containment_file": os.path.join(root, "containment_annotations.txt")}
    cfg = {"batch_size": batch_size, "inference_batch_size": 4, "num_workers": 0, "num_epochs": epochs, "print_step": 1,
           "learning_rate": 0.001, "lr_scheduler_patience": 2, "lr_scheduler_factor": 0.8, "device": device,
           "checkpoints_path": checkpoints}
    for split in ("train", "dev"):
        for k, v in data.items():
            cfg[f"{split}_{k}"] = v
    return cfg


def inference_config(root: str, device: str, model_path: str, batch_size: int = 2):
    return {"batch_size": batch_size, "num_workers": 0, "device": device, "model_path": model_path,
            "videos_dir": os.path.join(root, "videos"), "sample_dir": os.path.join(root, "od_perception"),
            "labels_dir": os.path.join(root, "labels")}
