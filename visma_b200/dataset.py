"""A stand-in for VISMA's clutter1 data (BASELINE config 2; the real recording is not available offline) in the
reference's own file layout, generated from the seeded synthetic room:

    <dataroot>/<dataset>/test.klg.ply            the RGB-D scene cloud (src/annotation.cpp:76, src/evaluation.cpp:124)
    <dataroot>/<dataset>/fragments/floor.ply     a patch of floor (src/annotation.cpp:79-80)
    <dataroot>/<dataset>/fragments/<model>_<k>.ply   the scene points of object k (:111)
    <dataroot>/<dataset>/fragments/objects.json  {"entries": [...]}  (:99-103)
    <dataroot>/<dataset>/fragments/alignment_gt.json   ground truth: key -> 12 numbers, 3x4 row-major
                                                 (the layout AnnotationTool writes to alignment.json, :153,175)
    <dataroot>/<dataset>/result.json             a semantic-mapping result in its own ("corvis") frame: the last
                                                 packet lists {id, status, model_name, model_pose[12]} (src/evaluation.cpp:162-175)
    <CAD_database_root>/<model>.obj              one CAD mesh per object (each object has its own scale)
    <root>/cfg/tool.json                         the reference's cfg/tool.json keys pointing at the above
"""
import os

import numpy as np

from . import io3d, synth


def write_clutter_dataset(root, n_scene=2_000_000, n_objects=8, seed=20260117, dataset="clutter1",
                          fragment_radius=0.75, pose_noise=(0.02, 0.01)):
    """Returns a dict with the paths and the ground truth (T_gt per key, T_ef_corvis)."""
    rng = np.random.default_rng(seed + 99)
    d = synth.make_room_scene(n_scene, n_objects, 16, seed=seed)
    scene_dir = os.path.join(root, "data", dataset)
    frag_dir = os.path.join(scene_dir, "fragments")
    cad_dir = os.path.join(root, "CAD")
    for p in (frag_dir, cad_dir, os.path.join(root, "cfg")):
        os.makedirs(p, exist_ok=True)
    xyz, nrm = d["scene_xyz"], d["scene_nrm"]
    io3d.write_ply(os.path.join(scene_dir, "test.klg.ply"), xyz, nrm)
    # floor fragment: floor points of one corner of the room
    floor = xyz[(np.abs(xyz[:, 1]) < 0.01) & (xyz[:, 0] < 1.0) & (xyz[:, 2] < 1.0)]
    io3d.write_ply(os.path.join(frag_dir, "floor.ply"), floor)
    V0, F = synth.load_chair()
    V0 = V0.astype(np.float64)
    V0[:, 1] -= V0[:, 1].min()
    entries, gt = [], {}
    for b in range(n_objects):
        # the same per-object scale make_room_scene drew (first draw of its per-object generator)
        scale = np.random.default_rng(seed + 1 + b).uniform(0.8, 1.2)
        model = "aeron%d" % b
        key = "%s_0" % model
        io3d.write_obj(os.path.join(cad_dir, model + ".obj"), V0 * scale, F)
        T = d["T_gt"][b]
        c = T[:3, 3]
        # the object's fragment: scene points above the floor around the object's position
        keep = (np.linalg.norm(xyz[:, [0, 2]] - c[[0, 2]], axis=1) < fragment_radius * scale) & (xyz[:, 1] > 0.015)
        io3d.write_ply(os.path.join(frag_dir, key + ".ply"), xyz[keep])
        entries.append(key)
        gt[key] = io3d.matrix_to_json(T[:3, :4])
    io3d.save_json({"entries": entries}, os.path.join(frag_dir, "objects.json"))
    io3d.save_json(gt, os.path.join(frag_dir, "alignment_gt.json"))
    # semantic-mapping result: the objects in another frame (an unknown rigid transform of the scene frame), each
    # pose a little off (what a visual-inertial mapper delivers)
    T_ef_corvis = synth.make_T(synth.rot_xyz(0.02, 0.7, -0.015), [0.4, -0.05, 0.9])
    inv = np.linalg.inv(T_ef_corvis)
    packet = []
    for b in range(n_objects):
        da = rng.normal(0, pose_noise[1], 3)
        dT = synth.make_T(synth.rot_xyz(*da), rng.normal(0, pose_noise[0], 3))
        pose = inv @ d["T_gt"][b] @ dT
        packet.append({"id": b, "status": 1, "model_name": "aeron%d" % b, "model_pose": io3d.matrix_to_json(pose[:3, :4])})
    io3d.save_json([[], packet], os.path.join(scene_dir, "result.json"))
    cfg = {"dataroot": os.path.join(root, "data") + "/", "dataset": dataset, "CAD_database_root": cad_dir + "/",
           "experiment_root": root, "datatype": "VLSLAM", "debug": False,
           "visualization": {"model_samples": 5000},
           "ICP": {"voxel_size": 0.01, "point_to_plane": False, "rotation_level": 24, "distance_threshold": 0.02},
           "evaluation": {"show_annotation": False, "ICP_refinement": True, "use_point_to_plane": False,
                          "voxel_size": 0.05, "max_distance": 0.075, "samples_per_model": 50000}}
    cfg_path = os.path.join(root, "cfg", "tool.json")
    with open(cfg_path, "w") as f:  # with a // comment, as the reference's own cfg/tool.json has
        f.write("{\n  // generated stand-in for VISMA's clutter1 (visma_b200/dataset.py)\n")
        import json
        f.write(json.dumps(cfg, indent=2)[1:])
    return dict(cfg_path=cfg_path, cfg=cfg, scene_dir=scene_dir, fragment_dir=frag_dir, cad_dir=cad_dir,
                entries=entries, T_gt={k: io3d.matrix_from_json(v, 3, 4) for k, v in gt.items()},
                T_ef_corvis=T_ef_corvis)
