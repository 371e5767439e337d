"""ORACLE / CPU BASELINE (test infrastructure, never shipped or measured as the product).

Runs the REFERENCE's own host code — deepsignal/call_modifications.py `call_mods` (:417-495) with its
reader / caller / writer processes (`_read_features_file` :35-91, `_call_mods_q` :197-234 -> `_call_mods`
:149-194, `_write_predstr_to_file` :283-297) — unmodified, from oracle/_ref/ (see oracle/make_ref.py) or
/root/reference.  TensorFlow 1.x cannot be installed in this image, so the three TF touch points of that
file are stubbed:

  * `Model(...)`            -> a handle whose attributes are the feed_dict keys,
  * `tf.Session().run(...)` -> the torch-CPU fp32 restatement of the UN-folded graph (oracle/model_oracle.py),
  * `tf.train.Saver().restore(sess, path)` -> loads the .npz checkpoint written by deepsignal_b200.weights.

Everything else (text parsing, batching, probability normalisation, str(np.float32) formatting, the process
fan-out and its 1-second polling sleeps) is the reference's code path as shipped.  Label: "TF-equivalent
CPU restatement", never "TF" (SURVEY.md section 8d).
"""
from __future__ import annotations

import contextlib
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

_state = {"ckpt_path": None, "threads": 1}


def ref_root():
    for cand in (os.path.join(HERE, "_ref"), os.environ.get("DEEPSIGNAL_REFERENCE", "/root/reference")):
        if cand and os.path.isdir(os.path.join(cand, "deepsignal")):
            return cand
    return None


class _FakeSession(object):
    """Stands in for tf.Session: run() evaluates the un-folded graph with torch on the CPU (fp32)."""

    def __init__(self, *a, **k):
        import torch
        torch.set_num_threads(int(_state["threads"]))
        self.ck = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def load(self, path):
        from deepsignal_b200 import weights as W
        self.ck = W.load_checkpoint(path)

    def run(self, fetches, feed_dict=None):
        import torch
        from oracle import model_oracle as mo
        if feed_dict is None:          # sess.run(tf.local_variables_initializer())
            return None
        fd = {getattr(k, "name", k): v for k, v in feed_dict.items()}
        probs, pred = mo.forward(self.ck, np.asarray(fd["base_int"], np.int64), np.asarray(fd["means"], np.float32),
                                 np.asarray(fd["stds"], np.float32), np.asarray(fd["sanums"], np.float32),
                                 np.asarray(fd["signals"], np.float32), dtype=torch.float32)
        return probs.astype(np.float32), pred.astype(np.int64)


class _Placeholder(object):
    def __init__(self, name):
        self.name = name

    def __hash__(self):
        return hash(self.name)

    def __eq__(self, other):
        return isinstance(other, _Placeholder) and other.name == self.name


class _FakeModel(object):
    """Attribute-compatible with deepsignal/model.py:25-56 as far as call_modifications.py touches it."""

    def __init__(self, base_num, signal_num, class_num, is_cnn=True, is_rnn=True, is_base=True, **kw):
        for n in ("base_int", "means", "stds", "sanums", "signals", "labels", "lr", "training", "keep_prob",
                  "activation_logits", "prediction"):
            setattr(self, n, _Placeholder(n))


class _FakeSaver(object):
    def restore(self, sess, path):
        sess.load(path)


def _install_stubs():
    np.int = int
    np.float = float
    h5py = types.ModuleType("h5py")
    sm = types.ModuleType("statsmodels")
    robust = types.ModuleType("statsmodels.robust")
    robust.mad = lambda a: np.median(np.fabs(np.asarray(a) - np.median(a)) / 0.6744897501960817)
    sm.robust = robust
    tf = types.ModuleType("tensorflow")
    contrib = types.ModuleType("tensorflow.contrib")
    rnn = types.ModuleType("tensorflow.contrib.rnn")
    rnn.LSTMCell = rnn.GRUCell = object
    layers = types.ModuleType("tensorflow.contrib.layers")
    layers.batch_norm = object
    framework = types.ModuleType("tensorflow.contrib.framework")
    contrib.rnn, contrib.layers, contrib.framework = rnn, layers, framework
    tf.contrib = contrib
    tf.float32 = "float32"
    tf.Session = _FakeSession
    tf.ConfigProto = lambda *a, **k: types.SimpleNamespace(gpu_options=types.SimpleNamespace(allow_growth=False))
    tf.train = types.SimpleNamespace(Saver=_FakeSaver)
    tf.local_variables_initializer = lambda: None
    for name, mod in [("h5py", h5py), ("statsmodels", sm), ("statsmodels.robust", robust), ("tensorflow", tf),
                      ("tensorflow.contrib", contrib), ("tensorflow.contrib.rnn", rnn),
                      ("tensorflow.contrib.layers", layers), ("tensorflow.contrib.framework", framework)]:
        sys.modules[name] = mod


_ref_cm = None


def load_reference():
    """Import the reference's deepsignal.call_modifications (unmodified source) with the TF stubs in place."""
    global _ref_cm
    if _ref_cm is not None:
        return _ref_cm
    root = ref_root()
    if root is None:
        raise ImportError("the reference package is neither under oracle/_ref/ (python oracle/make_ref.py) nor /root/reference")
    _install_stubs()
    mine = sys.modules.pop("deepsignal", None)      # make sure `deepsignal` resolves to the reference, not to a namesake
    sys.path.insert(0, root)
    try:
        import deepsignal.call_modifications as cm
    finally:
        sys.path.remove(root)
        if mine is not None:
            sys.modules["deepsignal"] = mine
    cm.Model = _FakeModel
    _ref_cm = cm
    return cm


def f5_args_default():
    # (is_recursive, corrected_group, basecall_subgroup, reference_path, is_dna, normalize_method, motifs, mod_loc,
    #  methy_label, f5_batch_num, position_file) -- call_modifications.py:428-429; only f5_batch_num matters for a feature file
    return (True, "RawGenomeCorrected_000", "BaseCalled_template", None, True, "mad", "CG", 0, 1, 20, None)


def run_call_mods(tsv_path, ckpt_path, out_path, nproc, torch_threads, batch_size=512):
    """reference `call_mods(..., is_gpu=False)` on a feature file; returns wall seconds of the whole call."""
    cm = load_reference()
    _state["ckpt_path"] = ckpt_path
    _state["threads"] = int(torch_threads)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(sys.stderr):       # the reference prints progress lines; keep stdout for the JSON line
        cm.call_mods(tsv_path, ckpt_path, out_path, 17, 360, batch_size, 0.001, 2, int(nproc), False, True, True, True,
                     f5_args_default())
    return time.perf_counter() - t0


def run_call_mods_guarded(tsv_path, ckpt_path, out_path, nproc, torch_threads, timeout_s):
    """`run_call_mods` in a fresh interpreter and its own process group, killed after ``timeout_s``.

    The reference can deadlock with several caller processes: its Queue subclass (utils/process_utils.py:507-537)
    decrements its shared size counter BEFORE the blocking get() and tests `empty()` as `counter == 0`, so two callers
    that race for the last item drive the counter negative and one of them blocks in get() forever — call_mods then
    never returns from `p.join()` (call_modifications.py:481-482).  A watchdog is therefore part of timing it.
    Returns wall seconds, or None if the run had to be killed."""
    import json
    import signal
    import subprocess
    cmd = [sys.executable, "-m", "oracle.ref_host", tsv_path, ckpt_path, out_path, str(int(nproc)), str(int(torch_threads))]
    proc = subprocess.Popen(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, start_new_session=True)
    try:
        out, _ = proc.communicate(timeout=timeout_s)
    except subprocess.TimeoutExpired:
        try:
            os.killpg(proc.pid, signal.SIGKILL)
        except ProcessLookupError:
            pass
        proc.wait()
        return None
    finally:
        try:                        # daemonic children of a finished run are gone; make sure none survives a failed one
            os.killpg(proc.pid, signal.SIGKILL)
        except (ProcessLookupError, PermissionError):
            pass
    for line in reversed(out.strip().splitlines()):
        if line.startswith("{"):
            return float(json.loads(line)["seconds"])
    return None


def time_layouts(tsv_path, n_lines, ckpt_path, out_dir, cores=None, probe_only=None):
    """Both process layouts of SURVEY 8d: one caller with all threads (nproc=3 -> nproc_tf=1) and cores-2
    single-thread callers (nproc=cores, call_modifications.py:461-476).  Returns {layout: samples/s}; a layout whose
    run deadlocked (see run_call_mods_guarded) reports 0.0."""
    cores = cores or os.cpu_count() or 1
    res = {}
    layouts = {"1 caller x {} threads".format(cores): (3, cores),
               "{} callers x 1 thread".format(max(1, cores - 2)): (max(3, cores), 1)}
    for name, (nproc, thr) in layouts.items():
        if probe_only and name not in probe_only:
            continue
        dt = run_call_mods_guarded(tsv_path, ckpt_path, os.path.join(out_dir, "ref_calls.tsv"), nproc, thr,
                                   timeout_s=90.0 + n_lines / 150.0)
        res[name] = (n_lines / dt) if dt else 0.0
    return res


if __name__ == "__main__":
    import json
    _tsv, _ck, _out, _nproc, _thr = sys.argv[1:6]
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    print(json.dumps({"seconds": run_call_mods(_tsv, _ck, _out, int(_nproc), int(_thr))}))
