#! /usr/bin/env python
"""`deepsignal` command line of the B200 engine: the ``extract`` and ``call_mods`` modules with the
reference's flags and defaults (reference deepsignal/deepsignal.py:156-326).  ``train`` / ``denoise``
are outside this implementation's scope (the call_mods hot path) and are not provided.

Extra flags: ``--precision {fp32,fp16x3,fp16x2,fp16,bf16}`` (call_mods; default fp16x2 = split-operand tensor-core path) and
``--device`` (CUDA device index).
"""
from __future__ import absolute_import

import argparse
import sys

from .utils.process_utils import str2bool, display_args


def main_extraction(args):
    from .extract_features import extract_features
    display_args(args)
    extract_features(args.fast5_dir, str2bool(args.recursively), args.reference_path, str2bool(args.is_dna),
                     args.f5_batch_num, args.write_path, args.nproc, args.corrected_group, args.basecall_subgroup,
                     args.normalize_method, args.motifs, args.mod_loc, args.kmer_len, args.cent_signals_len,
                     args.methy_label, args.positions, str2bool(args.w_is_dir), args.w_batch_num, device=args.device)


def main_call_mods(args):
    from .call_modifications import call_mods
    display_args(args)
    methy_label = 1   # reference deepsignal.py:79
    f5_args = (str2bool(args.recursively), args.corrected_group, args.basecall_subgroup, args.reference_path,
               str2bool(args.is_dna), args.normalize_method, args.motifs, args.mod_loc, methy_label,
               args.f5_batch_num, args.positions)
    call_mods(args.input_path, args.model_path, args.result_file, args.kmer_len, args.cent_signals_len,
              args.batch_size, args.learning_rate, args.class_num, args.nproc, str2bool(args.is_gpu),
              str2bool(args.is_rnn), str2bool(args.is_base), str2bool(args.is_cnn), f5_args,
              precision=args.precision, device=args.device,
              devices=[int(d) for d in args.devices.split(",")] if args.devices else None)


def _fast5_flags(g, with_label):
    g.add_argument("--recursively", "-r", action="store", type=str, required=False, default="yes",
                   help="is to find fast5 files from fast5 dir recursively. default true, t, yes, 1")
    g.add_argument("--corrected_group", action="store", type=str, required=False, default="RawGenomeCorrected_000",
                   help="the corrected_group of fast5 files after tombo re-squiggle. default RawGenomeCorrected_000")
    g.add_argument("--basecall_subgroup", action="store", type=str, required=False, default="BaseCalled_template",
                   help="the corrected subgroup of fast5 files. default BaseCalled_template")
    g.add_argument("--is_dna", action="store", type=str, required=False, default="yes",
                   help="whether the fast5 files from DNA sample or not. default true, t, yes, 1.")
    g.add_argument("--reference_path", action="store", type=str, required=False, default=None,
                   help="the reference file to be used, usually is a .fa file")
    g.add_argument("--normalize_method", action="store", type=str, choices=["mad", "zscore"], default="mad",
                   required=False, help="the way for normalizing signals in read level. mad or zscore, default mad")
    if with_label:
        g.add_argument("--methy_label", action="store", type=int, choices=[1, 0], required=False, default=1,
                       help="the label of the interested modified bases, this is for training. 0 or 1, default 1")
    g.add_argument("--motifs", action="store", type=str, required=False, default="CG",
                   help="motif seq to be extracted, default: CG. can be multi motifs splited by comma "
                        "(no space allowed in the input str), or use IUPAC alphabet, "
                        "the mod_loc of all motifs must be the same")
    g.add_argument("--mod_loc", action="store", type=int, required=False, default=0,
                   help="0-based location of the targeted base in the motif, default 0")
    g.add_argument("--positions", action="store", type=str, required=False, default=None,
                   help="file with a list of positions interested (tab-separated chromosome, position "
                        "(in fwd strand), strand). default None")


def main():
    parser = argparse.ArgumentParser(prog="deepsignal",
                                     description="detecting base modifications from Nanopore sequencing reads "
                                                 "(B200 engine): extract, call_mods",
                                     formatter_class=argparse.RawTextHelpFormatter)
    sub = parser.add_subparsers(title="modules", help="deepsignal modules, use -h/--help for help")
    se = sub.add_parser("extract", description="extract features from corrected (tombo) fast5s / packed reads")
    sc = sub.add_parser("call_mods", description="call modifications")

    gi = se.add_argument_group("INPUT")
    gi.add_argument("--fast5_dir", "-i", action="store", type=str, required=True,
                    help="the directory of fast5 files (or a packed-reads .npz / directory of them)")
    ge = se.add_argument_group("EXTRACTION")
    _fast5_flags(ge, with_label=True)
    ge.add_argument("--kmer_len", action="store", type=int, required=False, default=17, help="len of kmer. default 17")
    ge.add_argument("--cent_signals_len", action="store", type=int, required=False, default=360,
                    help="the number of signals to be used in deepsignal, default 360")
    go = se.add_argument_group("OUTPUT")
    go.add_argument("--write_path", "-o", action="store", type=str, required=True,
                    help="file path to save the features")
    go.add_argument("--w_is_dir", action="store", type=str, required=False, default="no",
                    help="if using a dir to save features into multiple files")
    go.add_argument("--w_batch_num", action="store", type=int, required=False, default=200,
                    help="features batch num to save in a single writed file when --is_dir is true")
    se.add_argument("--nproc", "-p", action="store", type=int, default=1, required=False,
                    help="number of processes to be used, default 1")
    se.add_argument("--f5_batch_num", action="store", type=int, default=50, required=False,
                    help="number of files to be processed by each process one time, default 50")
    se.add_argument("--device", type=int, default=0, help="CUDA device index")
    se.set_defaults(func=main_extraction)

    ci = sc.add_argument_group("INPUT")
    ci.add_argument("--input_path", "-i", action="store", type=str, required=True,
                    help="the input path, can be a signal_feature file from extract_features.py, "
                         "or a directory of fast5 files / packed reads")
    ci.add_argument("--f5_batch_num", action="store", type=int, default=50, required=False,
                    help="number of reads/files to be processed by each process one time, default 50")
    cc = sc.add_argument_group("CALL")
    cc.add_argument("--model_path", "-m", action="store", type=str, required=True,
                    help="file path of the trained model (.ckpt TF bundle prefix or .npz)")
    cc.add_argument("--is_cnn", type=str, default="yes", required=False, help="dose the used model contain inception module?")
    cc.add_argument("--is_rnn", type=str, default="yes", required=False, help="dose the used model contain BiLSTM module?")
    cc.add_argument("--is_base", type=str, default="yes", required=False,
                    help="dose the BiLSTM module of the used model take base features as input?")
    cc.add_argument("--kmer_len", "-x", action="store", default=17, type=int, required=False,
                    help="base num of the kmer, default 17")
    cc.add_argument("--cent_signals_len", "-y", action="store", default=360, type=int, required=False,
                    help="the number of central signals of the kmer to be used, default 360")
    cc.add_argument("--batch_size", "-b", default=512, type=int, required=False, action="store",
                    help="batch size, default 512")
    cc.add_argument("--learning_rate", "-l", default=0.001, type=float, required=False, action="store",
                    help="init learning rate, default 0.001")
    cc.add_argument("--class_num", "-c", action="store", default=2, type=int, required=False, help="class num, default 2")
    cc.add_argument("--precision", default="fp16x2", choices=["fp32", "fp16x3", "fp16x2", "fp16", "bf16"],
                    help="arithmetic path: fp32 CUDA cores, or tcgen05 tensor cores with split fp16 hi+lo activations "
                         "(fp16x2: within 2e-3 of the fp32 reference on the benchmarked checkpoint; fp16x3: weights split too, ~1e-5 whatever the "
                         "checkpoint, about half the rate), or single fp16 / bf16 operands (faster, looser)")
    co = sc.add_argument_group("OUTPUT")
    co.add_argument("--result_file", "-o", action="store", type=str, required=True,
                    help="the file path to save the predicted result")
    _fast5_flags(sc.add_argument_group("FAST5_EXTRACTION"), with_label=False)
    sc.add_argument("--nproc", "-p", action="store", type=int, default=1, required=False,
                    help="number of processes to be used, default 1.")
    sc.add_argument("--is_gpu", action="store", type=str, default="no", required=False, choices=["yes", "no"],
                    help="accepted for compatibility: this engine always runs on the GPU")
    sc.add_argument("--device", type=int, default=0, help="CUDA device index")
    sc.add_argument("--devices", type=str, default=None,
                    help="comma-separated CUDA device indices: shard the reads over several GPUs of this box "
                         "(one worker + engine per device, no collective; the result file is identical to a 1-GPU run). "
                         "Under torchrun every rank uses cuda:LOCAL_RANK instead.")
    sc.set_defaults(func=main_call_mods)

    args = parser.parse_args()
    if hasattr(args, "func"):
        args.func(args)
    else:
        parser.print_help()
    return 0


if __name__ == "__main__":
    sys.exit(main())
