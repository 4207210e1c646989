"""CPU stand-in with the reference estimator's interface, used ONLY to exercise the script harness of
tests/test_reference_scripts.py on a box without a GPU.  TEST INFRASTRUCTURE: the product never imports it."""
import time

import numpy as np

from oracle import prepost
from oracle.forward import OracleNet
from oracle.weights import make_weights


class VNectEstimator:
    box_size = 368
    hm_factor = 8
    joints_sum = 21
    joint_parents = [16, 15, 1, 2, 3, 1, 5, 6, 14, 8, 9, 14, 11, 12, 14, 14, 1, 4, 7, 10, 13]

    def __init__(self):
        print('Initializing VNect Estimator...')
        self.scales = [1, 0.85, 0.7]
        self._net = OracleNet(make_weights("W0"))
        self._est = None
        print('VNect Estimator initialized.')

    def __call__(self, img):
        t0 = time.time()
        if self._est is None or self._est.scales != list(self.scales):
            self._est = prepost.OracleEstimator(self._net, self.scales, clock=lambda: time.time())
        j2, j3 = self._est(np.ascontiguousarray(img))
        print('FPS: {:>2.2f}'.format(1 / (time.time() - t0)))
        return j2, j3
