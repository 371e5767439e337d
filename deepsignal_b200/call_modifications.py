"""call modifications from extracted features (or packed reads) with the B200 engine.

Host-side mirror of the reference's ``deepsignal/call_modifications.py``: same function
names, argument meaning, feature-file parsing (:35-91), per-sample output formatting
(:183-190) and 10-column result TSV (chrom, pos, strand, pos_in_strand, read_name,
read_strand, prob_0, prob_1, called_label, k_mer).  The TensorFlow session of the
reference (``sess.run([activation_logits, prediction], feed_dict)``, :168-178) is replaced
by one call into libdsb200 (``Engine.forward`` -> ``dsb_forward_host``); nothing on the
hot path is computed in Python.

Differences that are deliberate:
* one reader thread + one GPU worker per device instead of N TensorFlow processes
  (the reference's process fan-out exists to use CPU cores; the GPU engine is fed by
  large batches instead).  ``--nproc`` is accepted and only bounds host parsing threads.
* a features batch is forwarded in one call rather than in ``batch_size`` slices: the model is
  in inference mode, so results are batch-composition invariant; ``batch_num`` is still
  reported as ceil(n / batch_size) like the reference.
"""
from __future__ import absolute_import

import os
import sys
import threading
import time
import queue

import numpy as np

from .utils.process_utils import base2code_dna, code2base_dna, str2bool, display_args  # noqa: F401
from .utils.process_utils import get_motif_seqs, get_contig2len

queen_size_border = 100


def _parse_feature_line(words):
    return ("\t".join(words[0:6]),
            [base2code_dna[x] for x in words[6]],          # KeyError on non-ACGTN like the reference (:52)
            [float(x) for x in words[7].split(",")],
            [float(x) for x in words[8].split(",")],
            [int(x) for x in words[9].split(",")],
            [float(x) for x in words[10].split(",")],
            int(words[11]))


def _read_features_file(features_file, features_batch_q, f5_batch_num=20):
    """Group feature-file lines into batches of ``f5_batch_num`` *reads* (reference :35-91).

    Puts 7-tuples (sampleinfo, kmers, base_means, base_stds, base_signal_lens, cent_signals,
    labels) on the queue and finally the string "kill"."""
    r_num = 0
    cols = ([], [], [], [], [], [], [])
    with open(features_file, "r") as rf:
        readid_pre = None
        for line in rf:
            words = line.strip().split("\t")
            if readid_pre is None:
                readid_pre = words[4]
            elif words[4] != readid_pre:
                r_num += 1
                readid_pre = words[4]
                if r_num % f5_batch_num == 0:
                    features_batch_q.put(cols)
                    cols = ([], [], [], [], [], [], [])
            for dst, v in zip(cols, _parse_feature_line(words)):
                dst.append(v)
    if len(cols[0]) > 0:
        features_batch_q.put(cols)
    features_batch_q.put("kill")


def _call_mods(features_batch, engine, model, init_learning_rate, batch_size):
    """Mirror of reference ``_call_mods`` (:149-194): returns (pred_str, accuracy, batch_num).

    ``engine`` plays the role of the TF session, ``model`` carries the precision ("fp32", "fp16x2",
    "fp16", "bf16").  ``init_learning_rate`` is accepted for signature compatibility and unused (the
    reference only feeds it to a placeholder)."""
    sampleinfo, kmers, base_means, base_stds, base_signal_lens, cent_signals, labels = features_batch
    labels = np.reshape(labels, (len(labels)))
    n = len(sampleinfo)
    if n == 0:
        return [], float("nan"), 0
    kmers_a = np.asarray(kmers, dtype=np.int32)
    probs, prediction = engine.forward(kmers_a, np.asarray(base_means, np.float32), np.asarray(base_stds, np.float32),
                                       np.asarray(base_signal_lens, np.float32), np.asarray(cent_signals, np.float32),
                                       precision=getattr(model, "precision", "fp16x2"))
    prediction = prediction.astype(np.int64)
    batch_num = 0
    accuracys = []
    for i in range(0, n, batch_size):
        accuracys.append(float(np.mean(labels[i:i + batch_size] == prediction[i:i + batch_size])))
        batch_num += 1
    # prob_0/(prob_0+prob_1) in float32, printed with numpy's shortest repr — exactly reference :185-190
    p0, p1 = probs[:, 0], probs[:, 1]
    s = p0 + p1
    p0n, p1n = p0 / s, p1 / s
    pred_str = []
    for idx in range(n):
        pred_str.append("\t".join([sampleinfo[idx], str(p0n[idx]), str(p1n[idx]), str(prediction[idx]),
                                   "".join([code2base_dna[x] for x in kmers[idx]])]))
    return pred_str, float(np.mean(accuracys)), batch_num


def _write_predstr_to_file(write_fp, predstr_q):
    with open(write_fp, "w") as wf:
        while True:
            pred_str = predstr_q.get()
            if isinstance(pred_str, str) and pred_str == "kill":
                break
            for one_pred_str in pred_str:
                wf.write(one_pred_str + "\n")
            wf.flush()


class _ModelHandle(object):
    """What the reference's ``Model(base_num, signal_num, class_num, ...)`` is to ``_call_mods``."""

    def __init__(self, base_num, signal_num, class_num, is_cnn=True, is_base=True, is_rnn=True, precision="fp16x2"):
        if not (is_cnn or is_rnn):
            raise ValueError("at least one of is_cnn/is_rnn should be True")
        if class_num != 2:
            raise ValueError("the call_mods engine supports class_num == 2 (the reference default)")
        self.base_num, self.signal_num, self.class_num = base_num, signal_num, class_num
        self.is_cnn, self.is_base, self.is_rnn, self.precision = is_cnn, is_base, is_rnn, precision


def _load_engine(model_path, kmer_len, cent_signals_len, is_cnn, is_rnn, is_base, device=0):
    from .engine import Engine
    from .weights import load_checkpoint
    eng = Engine(device)
    eng.load_weights(load_checkpoint(model_path), is_cnn=is_cnn, is_rnn=is_rnn, is_base=is_base,
                     kmer_len=kmer_len, signal_len=cent_signals_len)
    return eng


def _features_batches_from_reads(reads_path, engine, f5_args, kmer_len, cent_signals_len):
    """fast5-directory mode of the reference (:94-122) on packed reads: extraction on the GPU,
    7-tuples out, grouped by ``f5_batch_num`` reads."""
    from .extract_features import load_reads, extract_batch_arrays
    (_, _, _, reference_path, is_dna, normalize_method, motifs, mod_loc, methy_label, f5_batch_num,
     position_file) = f5_args
    motif_seqs = get_motif_seqs(motifs, is_dna)
    chrom2len = get_contig2len(reference_path) if reference_path is not None else None
    positions = None
    if position_file is not None:
        from .extract_features import _read_position_file
        positions = _read_position_file(position_file)
    reads = load_reads(reads_path, chrom2len)
    step = max(1, int(f5_batch_num)) * 64
    for lo in range(0, reads.n_reads, step):
        rb = reads.slice(lo, min(reads.n_reads, lo + step))
        arr, info = extract_batch_arrays(engine, rb, motif_seqs, mod_loc, kmer_len, cent_signals_len,
                                         normalize_method, positions)
        labels = [methy_label] * len(info)
        yield (info, arr["kmer"].tolist(), arr["means64"], arr["stds64"], arr["lens_i32"], arr["signals64"], labels)


def call_mods(input_path, model_path, result_file, kmer_len, cent_signals_len,
              batch_size, learning_rate, class_num, nproc, is_gpu, is_rnn, is_base, is_cnn,
              f5_args, precision="fp16x2", device=0):
    """Same positional signature as the reference ``call_mods`` (:417-495) plus ``precision`` /
    ``device``.  ``is_gpu`` is accepted for compatibility; this implementation always runs on the
    GPU and fails loudly if none is present."""
    start = time.time()
    model_path = os.path.abspath(model_path)
    input_path = os.path.abspath(input_path)
    engine = _load_engine(model_path, kmer_len, cent_signals_len, is_cnn, is_rnn, is_base, device)
    model = _ModelHandle(kmer_len, cent_signals_len, class_num, is_cnn, is_base, is_rnn, precision)
    f5_batch_num = f5_args[9]

    pred_str_q = queue.Queue(maxsize=queen_size_border)
    p_w = threading.Thread(target=_write_predstr_to_file, args=(result_file, pred_str_q), daemon=True)
    p_w.start()
    n_batches = 0
    if os.path.isdir(input_path) or input_path.endswith(".npz"):
        for features_batch in _features_batches_from_reads(input_path, engine, f5_args, kmer_len, cent_signals_len):
            pred_str, _, nb = _call_mods(features_batch, engine, model, learning_rate, batch_size)
            pred_str_q.put(pred_str)
            n_batches += nb
    elif input_path.endswith(".bin"):
        # the reference's 1628-byte binary sample records (scripts/generate_binary_feature_file.py): no read / position
        # columns exist in that format, so the six info columns are printed as '.'
        from .binary_format import read_binary
        rec = read_binary(input_path, kmer_len, cent_signals_len)
        step = 1 << 18
        for lo in range(0, len(rec["labels"]), step):
            sl = slice(lo, lo + step)
            n_here = len(rec["labels"][sl])
            batch = (["\t".join(["."] * 6)] * n_here, rec["kmer"][sl].tolist(), rec["means"][sl], rec["stds"][sl],
                     rec["lens"][sl], rec["signals"][sl], rec["labels"][sl])
            pred_str, _, nb = _call_mods(batch, engine, model, learning_rate, batch_size)
            pred_str_q.put(pred_str)
            n_batches += nb
    elif os.environ.get("DSB_PY_TSV", "0") != "1":
        # fast path: C++ text codec (csrc/tsv.cpp); identical parsing rules and byte-identical result lines
        from .fast_tsv import iter_feature_blocks, format_calls
        for block in iter_feature_blocks(input_path, kmer_len, cent_signals_len, nthreads=max(1, nproc)):
            probs, prediction = engine.forward(block.kmer, block.means, block.stds, block.lens, block.signals,
                                               precision=model.precision)
            pred_str_q.put([format_calls(block, probs, prediction).decode().rstrip("\n")])
            n_batches += -(-block.n // batch_size)
    else:
        features_batch_q = queue.Queue(maxsize=queen_size_border)
        p_rf = threading.Thread(target=_read_features_file, args=(input_path, features_batch_q, f5_batch_num * 64),
                                daemon=True)
        p_rf.start()
        while True:
            features_batch = features_batch_q.get()
            if isinstance(features_batch, str) and features_batch == "kill":
                break
            pred_str, _, nb = _call_mods(features_batch, engine, model, learning_rate, batch_size)
            pred_str_q.put(pred_str)
            n_batches += nb
        p_rf.join()
    pred_str_q.put("kill")
    p_w.join()
    engine.close()
    print("call_mods process {} ending, proceed {} batches".format(os.getpid(), n_batches))
    print("call_mods costs %.2f seconds.." % (time.time() - start))


def main():
    from .deepsignal import main as cli_main
    sys.argv.insert(1, "call_mods")
    return cli_main()


if __name__ == "__main__":
    sys.exit(main())
