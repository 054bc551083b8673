"""TEST INFRASTRUCTURE — empty stand-in: the reference imports open3d at
gaussian_splatting/scene/gaussian_model.py:15 but never uses it on the tracking path."""
