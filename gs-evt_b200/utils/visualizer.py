"""Optional visual output (overlay PNGs).  Cosmetic, off the timed path; needs cv2."""
import numpy as np


def _to_img(frame):
    a = frame.detach().float().cpu().numpy().squeeze()
    m = np.abs(a).max()
    a = a / m if m > 0 else a
    img = np.zeros(a.shape + (3,), np.uint8)
    img[..., 2] = (np.clip(a, 0, 1) * 255).astype(np.uint8)
    img[..., 0] = (np.clip(-a, 0, 1) * 255).astype(np.uint8)
    return img


def get_delta_Ie_img(delta_Ie):
    return _to_img(delta_Ie)


def get_delta_Ir_img(delta_Ir):
    return _to_img(delta_Ir)


def overlay_two_imgs(a, b, alpha=0.5):
    return (a.astype(np.float32) * alpha + b.astype(np.float32) * (1 - alpha)).astype(np.uint8)


def save_video(img_dir, out_path, fps=120):
    import os
    import cv2
    from gsevt.compat import natsorted
    files = natsorted([f for f in os.listdir(img_dir) if f.endswith(".png")])
    if not files:
        return
    first = cv2.imread(os.path.join(img_dir, files[0]))
    vw = cv2.VideoWriter(out_path, cv2.VideoWriter_fourcc(*"mp4v"), fps, (first.shape[1], first.shape[0]))
    for f in files:
        vw.write(cv2.imread(os.path.join(img_dir, f)))
    vw.release()


def save_gif(img_dir, out_path, duration=2):
    return None
