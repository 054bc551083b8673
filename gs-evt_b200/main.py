"""Entry point with the reference's CLI (main.py:18-56): python main.py -c <config.yaml>.
Run from this directory (gs-evt_b200/) or with it on PYTHONPATH."""
import os
import shutil
import sys
from argparse import ArgumentParser

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import torch
import yaml

from gsevt.compat import munchify
from utils.tracker import Tracker
from utils.render_camera.camera import Camera
from utils.event_camera.event import load_events_from_txt
from gaussian_splatting.scene.gaussian_model import GaussianModel


def main(config_path):
    with open(config_path, "r") as yml:
        config = yaml.safe_load(yml)
    model_params = munchify(config["Gaussian"]["model_params"])
    pipeline = munchify(config["Gaussian"]["pipeline_params"])
    device = model_params.device
    background = torch.tensor(model_params.background, dtype=torch.float32, device=device)
    os.makedirs(config["Tracking"]["save_path"], exist_ok=True)
    shutil.copy(config_path, config["Tracking"]["save_path"])

    viewpoint = Camera.init_from_yaml(config)
    gaussians = GaussianModel(model_params.sh_degree, device=device)
    gaussians.load_ply(model_params.model_path)
    event_arrays = load_events_from_txt(config["Event"]["data_path"], config["Event"]["max_events_per_frame"], array_nums=None)
    tracker = Tracker(config, event_arrays, viewpoint, gaussians, pipeline, background, device)
    tracker.tracking()
    return tracker


if __name__ == "__main__":
    parser = ArgumentParser(description="configuration parameters")
    parser.add_argument("--config_path", "-c", type=str, default="./configs/VECTOR/robot_normal1_config.yaml")
    args = parser.parse_args(sys.argv[1:])
    main(args.config_path)
