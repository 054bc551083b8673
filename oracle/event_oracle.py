"""TEST INFRASTRUCTURE — CPU oracle for the event side of the GS-EVT tracking path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  It is a numpy restatement of what the reference does with OpenCV on the host:

  E0  polarity scatter-add into a float32 (H, W) frame      reference utils/event_camera/event.py:118-120
  E1  cv2.undistort(frame, K, D)                            reference utils/event_camera/event.py:121
  E2  cv2.GaussianBlur(frame, (9, 9), 0, BORDER_REPLICATE)  reference utils/event_camera/event.py:122-123
  E3  cv2.normalize(frame, None)  (L2)                      reference utils/event_camera/event.py:124
  E4  abs                                                   reference utils/event_camera/event.py:126
  P   3-level INTER_NEAREST pyramid                         reference utils/tracker.py:78-91

OpenCV (pinned 4.8.1.78 in the reference's requirements.txt, 4.13.0 in this image) is a third-party
dependency whose source is not under /root/reference, so E1-E3 restate its published algorithm:
  * undistort == initUndistortRectifyMap(K, D, I, K, size, CV_16SC2) + remap(INTER_LINEAR,
    BORDER_CONSTANT 0): inverse map evaluated in float64, quantised to 1/32 pixel
    (INTER_BITS = 5), bilinear weights taken from the 32x32 float32 table;
  * GaussianBlur with sigma <= 0 and ksize 9 uses the fixed kernel [4,13,30,51,60,51,30,13,4]/256 (other odd sizes: the
    coefficients of cv2.getGaussianKernel(k, 0, CV_32F), tabulated in tests/golden/gauss_taps.npz by make_gauss_taps.py);
    row pass accumulates left to right with FMA, the column pass uses the symmetric form;
  * normalize(NORM_L2) multiplies by float32(1 / sqrt(sum_fp64 x^2)).
Pinned against cv2 itself in tests/test_event_oracle.py (cv2 is installed in the image, so the pin
also runs on the GPU box).
"""
import numpy as np

GAUSS9 = (np.array([4, 13, 30, 51, 60, 51, 30, 13, 4], dtype=np.float64) / 256.0).astype(np.float32)


def parse_events_txt(text):
    """reference utils/event_camera/event.py:20-22: each line is 'ts x y p' (ints)."""
    arr = np.array(text.split(), dtype=np.int64).reshape(-1, 4)
    return arr  # columns ts, x, y, p


def packetise(ev, max_events_per_frame, array_nums=None):
    """reference event.py:25-37: fixed-count packets, tail dropped."""
    n = ev.shape[0] // max_events_per_frame
    if array_nums is not None:
        n = min(n, array_nums)
    return [ev[i * max_events_per_frame:(i + 1) * max_events_per_frame] for i in range(n)]


def packet_duration(pkt):
    """reference event.py:92-97."""
    return (int(pkt[-1, 0]) - int(pkt[0, 0])) / 1e6


def packet_time(pkt):
    """reference event.py:99-100."""
    return (int(pkt[0, 0]) + (int(pkt[-1, 0]) - int(pkt[0, 0])) / 2) / 1e6


def accumulate(x, y, p, W, H):
    """E0 — integer-valued polarity histogram, exact in float32."""
    frame = np.zeros((H, W), dtype=np.int32)
    np.add.at(frame, (np.asarray(y, dtype=np.int64), np.asarray(x, dtype=np.int64)),
              np.where(np.asarray(p) != 0, 1, -1).astype(np.int32))
    return frame


def undistort_map(K, D, W, H):
    """Fixed-point (1/32 px) inverse map of cv2.undistort with newCameraMatrix = K, R = I."""
    K = np.asarray(K, dtype=np.float64).reshape(3, 3)
    D = np.asarray(D, dtype=np.float64).ravel()
    k1, k2, p1, p2 = D[0], D[1], D[2], D[3]
    k3 = D[4] if D.size > 4 else 0.0
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    u = np.arange(W, dtype=np.float64)[None, :]
    v = np.arange(H, dtype=np.float64)[:, None]
    x = (u - cx) / fx + 0 * v
    y = (v - cy) / fy + 0 * u
    x2, y2 = x * x, y * y
    r2 = x2 + y2
    _2xy = 2 * x * y
    kr = 1 + ((k3 * r2 + k2) * r2 + k1) * r2
    xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2)
    yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy
    mx = xd * fx + cx
    my = yd * fy + cy
    ix = np.rint(mx * 32).astype(np.int64)
    iy = np.rint(my * 32).astype(np.int64)
    return ix, iy


def undistort(frame, K, D):
    """E1 — bilinear remap with 5 fractional bits, constant-0 border (restates cv2.undistort)."""
    frame = np.asarray(frame, dtype=np.float32)
    H, W = frame.shape
    ix, iy = undistort_map(K, D, W, H)
    sx, sy = ix >> 5, iy >> 5
    fxb = (ix & 31).astype(np.float32) / np.float32(32)
    fyb = (iy & 31).astype(np.float32) / np.float32(32)

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        out = np.zeros((H, W), dtype=np.float32)
        out[ok] = frame[yy[ok], xx[ok]]
        return out

    one = np.float32(1)
    w00 = (one - fyb) * (one - fxb)
    w01 = (one - fyb) * fxb
    w10 = fyb * (one - fxb)
    w11 = fyb * fxb
    # integer counts times weights that are multiples of 1/1024: every product and sum is exact in
    # float32, so accumulation order does not matter here.
    acc = tap(sy, sx) * w00 + tap(sy, sx + 1) * w01 + tap(sy + 1, sx) * w10 + tap(sy + 1, sx + 1) * w11
    return acc.astype(np.float32)


def _fma(a, b, c):
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(np.float32)


def gaussian_blur9(frame):
    """E2 — separable fixed 9-tap kernel, replicate border, OpenCV's FMA accumulation order."""
    f = np.asarray(frame, dtype=np.float32)
    H, W = f.shape
    k = GAUSS9
    xp = np.pad(f, ((0, 0), (4, 4)), mode="edge")
    s = (xp[:, 0:W] * k[0]).astype(np.float32)
    for i in range(1, 9):
        s = _fma(xp[:, i:i + W], k[i], s)
    yp = np.pad(s, ((4, 4), (0, 0)), mode="edge")
    t = (yp[4:4 + H, :] * k[4]).astype(np.float32)
    for j in range(1, 5):
        pair = (yp[4 + j:4 + j + H, :] + yp[4 - j:4 - j + H, :]).astype(np.float32)
        t = _fma(pair, k[4 + j], t)
    return t


def gauss_taps(ksize):
    """Full k-tap kernel (float32) of cv2.getGaussianKernel(k, 0, CV_32F) from the committed table."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "gauss_taps.npz"))
    half = z["k%d" % int(ksize)].astype(np.float32)            # centre outwards
    return np.concatenate([half[:0:-1], half]).astype(np.float32)


def gaussian_blur(frame, ksize=9):
    """E2 for any odd kernel size 1..31: same pass structure as gaussian_blur9 (which it reproduces for ksize 9)."""
    f = np.asarray(frame, dtype=np.float32)
    H, W = f.shape
    k = gauss_taps(ksize)
    r = int(ksize) // 2
    xp = np.pad(f, ((0, 0), (r, r)), mode="edge")
    s = (xp[:, 0:W] * k[0]).astype(np.float32)
    for i in range(1, 2 * r + 1):
        s = _fma(xp[:, i:i + W], k[i], s)
    yp = np.pad(s, ((r, r), (0, 0)), mode="edge")
    t = (yp[r:r + H, :] * k[r]).astype(np.float32)
    for j in range(1, r + 1):
        pair = (yp[r + j:r + j + H, :] + yp[r - j:r - j + H, :]).astype(np.float32)
        t = _fma(pair, k[r + j], t)
    return t


def l2_normalize(frame):
    """E3 — cv2.normalize(x, None): x * float32(1 / ||x||_2) with the norm accumulated in float64."""
    f = np.asarray(frame, dtype=np.float32)
    nrm = np.sqrt(np.sum(f.astype(np.float64) ** 2))
    scale = np.float32(1.0 / nrm) if nrm > np.finfo(np.float64).eps else np.float32(0)
    return (f * scale).astype(np.float32)


def event_frame(x, y, p, W, H, K, D, ksize=9):
    """E0..E4 — returns (sign_delta_Ie, unsign_delta_Ie), each (1, H, W) float32."""
    cnt = accumulate(x, y, p, W, H).astype(np.float32)
    und = undistort(cnt, K, D)
    f = l2_normalize(gaussian_blur9(und) if int(ksize) == 9 else gaussian_blur(und, ksize))
    return f[None], np.abs(f)[None]


def pyramid(frame, levels=3):
    """P — cv2.resize(frame, (int(W s), int(H s)), INTER_NEAREST), s = 0.5**l (tracker.py:87-88; no re-normalisation).
    OpenCV's resizeNN samples source index min(floor(dst * (1 / (dst_size / src_size))), src_size - 1) in double: this is
    frame[::2**l, ::2**l] when the size is a multiple of 2**l and drifts from it otherwise (checked against cv2 itself in
    tests/test_event_oracle.py on 346 x 260)."""
    f = np.asarray(frame)
    H, W = f.shape[-2:]
    out = []
    for l in range(levels):
        Hl, Wl = int(H * 0.5 ** l), int(W * 0.5 ** l)
        iy = np.minimum(np.floor(np.arange(Hl) * (1.0 / (Hl / H))).astype(np.int64), H - 1)
        ix = np.minimum(np.floor(np.arange(Wl) * (1.0 / (Wl / W))).astype(np.int64), W - 1)
        out.append(np.ascontiguousarray(f[..., iy, :][..., ix]))
    return out
