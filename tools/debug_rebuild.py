"""Debug helper: filter-state export / import across an engine rebuild (GPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import prepost, synth
from oracle.forward import OracleNet
from oracle.weights import make_weights
from vnect_b200 import VNectEngine

w0 = make_weights("W0")
net = OracleNet(w0)
small = synth.stream_frame(5, 0)
big = np.random.default_rng(12).integers(0, 256, (400, 500, 3), dtype=np.uint8)
q = []
ref = prepost.OracleEstimator(net, [1.0], clock=lambda: q.pop(0))
q[:] = [50.0, 50.02]; r2a, r3a = ref(small)
q[:] = [50.04, 50.06]; r2b, r3b = ref(big)

# A: no rebuild
e = VNectEngine(w0, [1.0], max_frames=1, max_streams=1, max_input=(400, 500))
a2, a3 = e.estimate(small, [0], [50.0], [50.02]); a2 = a2.copy()
st = e.export_stream_state(0)
b2, b3 = e.estimate(big, [0], [50.04], [50.06])
print("A no rebuild: frame0 diff", np.abs(a2[0] - r2a).max(), "frame1 diff", np.abs(b2[0] - r2b).max())
st_after = e.export_stream_state(0)
e.close()
# B: rebuild + import
e = VNectEngine(w0, [1.0], max_frames=1, max_streams=1, max_input=(368, 368))
a2, a3 = e.estimate(small, [0], [50.0], [50.02])
st1 = e.export_stream_state(0)
print("export identical across engines:", np.array_equal(st, st1, equal_nan=True))
e.close()
e = VNectEngine(w0, [1.0], max_frames=1, max_streams=1, max_input=(400, 500))
e.import_stream_state(st1, 0)
st2 = e.export_stream_state(0)
print("import->export roundtrip:", np.array_equal(st1, st2, equal_nan=True), "first filter:", st1[:7], st2[:7])
b2, b3 = e.estimate(big, [0], [50.04], [50.06])
print("B rebuild: frame1 diff", np.abs(b2[0] - r2b).max(), "state after equal to A:", np.array_equal(e.export_stream_state(0), st_after, equal_nan=True))
e.close()
