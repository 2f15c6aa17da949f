"""One synthetic subject through the device part of the preprocess path (resample chest + wrist, window gather), twice; run under
ncu with --launch-skip <launches of the first pass> for the per-launch list of the second."""
import sys
import torch
sys.path.insert(0, ".")
from multimodalsignal_b200 import preprocess as pp, synth, _ext

lib = _ext.lib()
sub = synth.make_subject("S2", 0, minutes=100.0)
d = sub.as_pickle_dict()
chest = {k.decode(): v for k, v in d[b"signal"][b"chest"].items()}
wrist = {k.decode(): v for k, v in d[b"signal"][b"wrist"].items()}
dev = torch.device("cuda", 0)
up = pp._UPLOADER
staged = up.groups([up._columns(chest, pp.CHEST_CHANNELS)] + [up._columns(wrist, [n]) for n in pp.WRIST_CHANNELS], dev)
rows, wr = staged[0], dict(zip(pp.WRIST_CHANNELS, staged[1:]))
proto = pp.base_halving_quirk(sub.sid, sub.protocol)
torch.cuda.synchronize()
for rep in range(3):
    l0 = lib.mms_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    streams = pp.resample_subject_rows(rows, wr, 64)
    starts, labels, window = pp.window_plan(proto, 64)
    s = pp.SubjectStreams("S", streams, starts, labels, window, pp.CHEST_CHANNEL_NAMES + pp.WRIST_CHANNEL_NAMES)
    w = s.windows_f64()
    e1.record()
    torch.cuda.synchronize()
    print("rep", rep, "ms", round(e0.elapsed_time(e1), 3), "launches", lib.mms_launch_count() - l0, "windows", tuple(w.shape), flush=True)
