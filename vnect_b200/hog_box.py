"""Headless drop-in for the reference's ``HOGBox`` (src/hog_box.py:10-68): the bounding-box initialiser of the run scripts.

The reference shows every frame in an OpenCV window and waits for a mouse click before it accepts the current
detection (run_estimator.py:66-83); on a server there is no display and no mouse.  This class keeps the detector, the
rectangle arithmetic and the call contract -- ``choose, rect = hog(frame)`` with ``rect = [x, y, w, h]`` -- and replaces
the click by a rule: the box is accepted on the first frame that contains a person (or after ``give_up_after`` frames
without one, with the whole frame as the box, which is what the reference starts from, run_estimator.py:68).
``HOGBox.clicked = True`` before the first call accepts immediately, exactly like run_pic.py:21 does.

The detector itself (cv2.HOGDescriptor + the default people SVM) stays on the host: it runs once per video, not per
frame, and is not on the hot path.  The accepted rectangle seeds the on-device tracker
(``VNectEngine.set_box`` / vnect_track_set_box).
"""
import cv2
import numpy as np


class HOGBox:
    clicked = False  # src/hog_box.py:15 (class attribute; the scripts set it on the instance)

    def __init__(self, give_up_after=30, verbose=True):
        if verbose:
            print('Initializing HOGBox...')
        self.hog = cv2.HOGDescriptor()
        self.hog.setSVMDetector(cv2.HOGDescriptor_getDefaultPeopleDetector())
        self.give_up_after = int(give_up_after)
        self.frames_seen = 0
        if verbose:
            print('HOGBox initialized.')

    def __call__(self, img):
        """src/hog_box.py:25-38: largest detection, widened by cal_rect; whole frame when nothing is found.  The
        rectangle is drawn into ``img`` like the reference does (the scripts pass a copy or a throw-away frame)."""
        H, W = img.shape[:2]
        found, _ = self.hog.detectMultiScale(img)
        rect = self.cal_rect(found[np.argmax([found[i, 2] * found[i, 3] for i in range(len(found))])], H, W) \
            if len(found) else [0, 0, W, H]
        self.draw_rect(img, rect)
        self.frames_seen += 1
        if len(found) or self.frames_seen >= self.give_up_after:
            self.clicked = True  # stands in for on_mouse (src/hog_box.py:40-45)
        return self.clicked, rect

    @staticmethod
    def cal_rect(rect, H, W):
        """src/hog_box.py:47-58: widen the detection by 20 % of the frame width and 10 % of its height per side,
        clipped to the frame.  Same integer arithmetic, same return types (numpy integers in a list)."""
        x, y, w, h = rect
        offset_w = int(0.4 / 2 * W)
        offset_h = int(0.2 / 2 * H)
        x0, y0 = np.max([x - offset_w, 0]), np.max([y - offset_h, 0])
        return [x0, y0, np.min([x + w + offset_w, W]) - x0, np.min([y + h + offset_h, H]) - y0]

    @staticmethod
    def draw_rect(img, rect):
        """src/hog_box.py:60-68"""
        x, y, w, h = (int(v) for v in rect)
        cv2.rectangle(img, (x, y), (x + w, y + h), (60, 66, 207), 4)
