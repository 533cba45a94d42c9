"""The scalars the reference handlers read from nerf/configs/office_*_config.yaml (the four files
are byte-identical).  Kept as a dict so no YAML file has to ship; a caller
may pass the parsed reference YAML instead -- arithmetic strings such as "1024*32" are accepted."""
import copy

OFFICES = ("office_tokyo", "office_new_york", "office_geneve", "office_belgrade")

DEFAULT_CONFIG = {
    "experiment": {"image_width": 320, "image_height": 240, "endpoint_feat": False},
    "training": {"n_iterations": 200000, "learning_rate": 5e-4, "learning_rate_decay_rate": 0.1,
                 "learning_rate_decay_steps": 50000},
    "model": {"net_depth": 8, "net_width": 256, "net_depth_fine": 8, "net_width_fine": 256,
              "chunk": "1024*32", "net_chunk": "1024*32"},
    "rendering": {"n_rays": "32*32*1", "n_samples": 64, "n_importance": 128, "perturb": 1,
                  "use_view_dirs": True, "num_freqs_3d": 10, "num_freqs_2d": 4, "raw_noise_std": 1,
                  "test_viz_factor": 1, "depth_range": [0.1, 10.0], "white_background": False},
    "logging": {"step_log_print": 1, "step_log_tensorboard": 500, "step_save_ckpt": 20000,
                "step_render_test": 5000, "step_render_train": 5000},
    "inference": {"chunk": "1024*8"},
}


def default_config() -> dict:
    return copy.deepcopy(DEFAULT_CONFIG)


def number(value):
    """ints/floats pass through; strings are products of integers ("1024*32"), which is all the
    reference's eval() (inference handler:42-50) is ever used for."""
    if isinstance(value, str):
        out = 1
        for part in value.split("*"):
            out *= int(part.strip())
        return out
    return value
