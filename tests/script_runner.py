#!/usr/bin/env python3
"""Runs one of the reference's run scripts UNCHANGED inside a prepared working directory.

    python tests/script_runner.py <script.py> <workdir> <out.npz> <max_loop_iterations>

TEST INFRASTRUCTURE.  The working directory (made by tests/test_reference_scripts.py) holds a ``src`` package whose
``estimator.py`` is the binding under test, plus ``pic/``.  This runner only removes what a headless box cannot do:
  * OpenCV GUI calls (imshow / waitKey / namedWindow / setMouseCallback / destroy*) become no-ops; waitKey returns -1
    (no key) for the first ``max_loop_iterations`` calls and 27 afterwards, which is how the scripts' loops end;
  * ``cv2.VideoCapture`` serves the decoded frames of tests/golden/video.npz (the GPU box has no reference tree);
  * matplotlib-backed ``utils.draw_limbs_3d`` / ``plot_3d_init`` are replaced by recorders, ``utils.draw_limbs_2d`` is
    wrapped by one: what the script hands to the drawing code is what the test compares;
  * ``time.time`` is a scripted clock that logs (value, calling module), so the test knows the exact timestamps the
    estimator's OneEuroFilters saw.
The script text itself is exec'd as ``__main__`` without any edit.
"""
import os
import sys
import time
import types

import numpy as np


def main():
    script, workdir, out_path, max_iter = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
    golden_video = os.environ.get("VNECT_TEST_VIDEO_NPZ")
    os.chdir(workdir)
    sys.path.insert(0, workdir)

    # ---- scripted clock
    clock_log = []
    state = {"t": 1000.0}

    def fake_time():
        state["t"] += 0.01
        caller = sys._getframe(1).f_globals.get("__name__", "?")
        clock_log.append((state["t"], caller))
        return state["t"]
    time.time = fake_time

    # ---- matplotlib is not installed in this image: empty stand-ins (only touched by drawing code we replace)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "mpl_toolkits", "mpl_toolkits.mplot3d"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["matplotlib.animation"], "FuncAnimation"):
        sys.modules["matplotlib.animation"].FuncAnimation = object
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["mpl_toolkits"].mplot3d = sys.modules["mpl_toolkits.mplot3d"]

    # ---- OpenCV GUI / capture stubs
    import cv2
    gui = {"waitkey_calls": 0, "imshow": 0, "mouse_cb": None}

    def wait_key(delay=0):
        gui["waitkey_calls"] += 1
        return -1 if gui["waitkey_calls"] <= max_iter else 27

    def imshow(name, img):
        gui["imshow"] += 1
        if gui["mouse_cb"] is not None and gui["imshow"] == 2:  # the reference's HOGBox waits for a click: click once
            gui["mouse_cb"](cv2.EVENT_LBUTTONUP, 0, 0, 0, None)

    def set_mouse_callback(name, cb, *a):
        gui["mouse_cb"] = cb

    cv2.waitKey = wait_key
    cv2.imshow = imshow
    cv2.namedWindow = lambda *a, **k: None
    cv2.setMouseCallback = set_mouse_callback
    cv2.destroyWindow = lambda *a, **k: None
    cv2.destroyAllWindows = lambda *a, **k: None

    class FakeCapture:
        def __init__(self, source):
            self.frames = np.load(golden_video)["frames"] if golden_video else np.zeros((0, 2, 2, 3), np.uint8)
            self.i = 0

        def isOpened(self):
            return len(self.frames) > 0

        def get(self, prop):
            if prop == cv2.CAP_PROP_FRAME_WIDTH:
                return float(self.frames.shape[2])
            if prop == cv2.CAP_PROP_FRAME_HEIGHT:
                return float(self.frames.shape[1])
            return 0.0

        def read(self):
            if self.i >= len(self.frames):
                return False, None
            f = self.frames[self.i].copy()
            self.i += 1
            return True, f

        def release(self):
            pass
    cv2.VideoCapture = FakeCapture

    # ---- recorders around the drawing code
    from src import utils
    rec = {"j2": [], "rect": [], "j3": []}
    orig_2d = getattr(utils, "draw_limbs_2d", None)

    def draw_limbs_2d(img, joints_2d, limb_parents, rect):
        rec["j2"].append(np.array(joints_2d, dtype=np.float64))
        rec["rect"].append(np.array([int(v) for v in rect], dtype=np.int64))
        return orig_2d(img, joints_2d, limb_parents, rect) if orig_2d else img

    def draw_limbs_3d(joints_3d, joint_parents, *a, **k):
        rec["j3"].append(np.array(joints_3d))

    def plot_3d_init(joint_parents, joints_iter_gen, *a, **k):
        rec["gen"] = joints_iter_gen
    utils.draw_limbs_2d = draw_limbs_2d
    utils.draw_limbs_3d = draw_limbs_3d
    utils.plot_3d_init = plot_3d_init

    # ---- the script, unchanged
    with open(script) as f:
        text = f.read()
    glb = {"__name__": "__main__", "__file__": script}
    exec(compile(text, script, "exec"), glb)

    if "joints_3d" in glb and not rec["j3"]:
        rec["j3"].append(np.array(glb["joints_3d"]))  # run_estimator.py keeps the last one in a global for its animation
    est = glb.get("estimator")
    np.savez(out_path,
             j2=np.array(rec["j2"]), rect=np.array(rec["rect"]), j3=np.array(rec["j3"]),
             clock_t=np.array([t for t, _ in clock_log]), clock_who=np.array([w for _, w in clock_log]),
             estimator_class=str(type(est).__module__ + "." + type(est).__name__),
             final_rect=np.array([int(glb.get(k, -1)) for k in ("x", "y", "w", "h")]))


if __name__ == "__main__":
    main()
