"""``krisp_fasta`` command line on the B200 path — same flags, same output rows as the reference.

Mirrors ``krisp.krisp_fasta.krisp_fasta:main`` (krisp_fasta/krisp_fasta.py:126-298): argument
surface :128-176, (L, D, R, amplicon) deduction :178-213, stage sequence :236-291.  The three
file-to-file stages (sortedKmers*, mergeFiles, filterAlignments) are one device search; the host
renders the survivors.  ``--cores`` and ``--workdir`` are accepted and ignored (no worker processes,
no temporary k-mer files).  ``--primer3`` runs the reference's Primer3 post-filter on the survivors (host, ``render.run_primer3``);
it needs primer3-py, as the reference does.
"""
import argparse
import sys
import time

from .names import simplename
from .render import render_output
from .search import search_files


def extractSortedKmers(fasta, primer_left, primer_right, ampl_len, output, sortmem=None, parallel=1, verbose=True, omit=True):
    """Fasta file -> sorted ``left,mid,right`` k-mer table written to `output` — the reference's stage function
    (krisp_fasta/krisp_fasta.py:16-62) with the same arguments and the same progress lines on stderr; extraction and sort run on
    the device (``krisp_b200.kstream``: K1 + radix sort).  `sortmem` / `parallel` tuned GNU sort and are accepted for compatibility."""
    import time
    from .kstream import kstream
    kmers = kstream(fasta, kmers=ampl_len, disallow="Nn", complements=True, omitsoft=bool(omit), mapsoft=not omit,
                    split=[primer_left, -primer_right], sort=True, sortmem=sortmem, sortcols=[0, 2], sortnp=parallel, parallel=parallel)
    if not verbose:
        kmers.write(output)
        return
    start_t = time.time()
    print(f"Extracting {ampl_len}-mers from {fasta} and saving to {output}", end="\n", file=sys.stderr)
    found = kmers.write(output)
    print(f"=> Extracted and sorted {found:,} {ampl_len}-kmers from {fasta} in {time.time() - start_t:.2f}s", file=sys.stderr)


def build_parser():
    parser = argparse.ArgumentParser(
        description="Find diagnostic alignments for a set of fasta files",
        prog="krisp",
        formatter_class=argparse.RawTextHelpFormatter)
    parser.add_argument("files", nargs="+", type=str, metavar='PATH',
                        help="Fasta file to read. .gz, .bz2")
    parser.add_argument("--outgroup", nargs="*", type=str, default=[], metavar='PATH',
                        help="Outgroup Fasta files. To be amplified, but not detected")
    parser.add_argument("-c", "--conserved", type=int, metavar='INT',
                        help="Length of conserved regions on ends of amplicon")
    parser.add_argument("--conserved-left", type=int, metavar='INT',
                        help="Length of conserved region on left of amplicon")
    parser.add_argument("--conserved-right", type=int, metavar='INT',
                        help="Length of conserved region on right of amplicon")
    parser.add_argument("-d", "--diagnostic", type=int, metavar='INT',
                        help="Diagnostic region length for amplicon")
    parser.add_argument("-a", "--amplicon", type=int, metavar='INT',
                        help="Total amplicon length")
    parser.add_argument("--omit-soft", action="store_true",
                        help="Omit softmasked nucleotides")
    parser.add_argument("--cores", type=int, default=1, metavar='INT',
                        help="Accepted for compatibility; the search runs on the GPU. (default: %(default)s)")
    parser.add_argument("--dot-alignment", action="store_true",
                        help="Output as dot-based alignments")
    parser.add_argument("-o", "--out_align", type=str, metavar='PATH',
                        help="Write results as human-readable alignments to a file. (default: do not write alignment output)")
    parser.add_argument("-s", "--out_csv", type=str, metavar='PATH',
                        help="Write results to as a CSV (comma-separated value) file. (default: print to screen (stdout))")
    parser.add_argument("-w", "--workdir", type=str, metavar='PATH',
                        help="Accepted for compatibility; no temporary files are written")
    parser.add_argument("-p", "--primer3", action=argparse.BooleanOptionalAction,
                        help="Filter regions with Primer3 (needs primer3-py)")
    parser.add_argument('--tm', type=int, nargs=2, metavar='INT', default=[53, 68])
    parser.add_argument('--gc', type=int, nargs=2, metavar='INT', default=[40, 70])
    parser.add_argument('--amp_size', type=int, nargs=2, metavar='INT', default=[70, 150])
    parser.add_argument('--primer_size', type=int, nargs=2, metavar='INT', default=[25, 35])
    parser.add_argument('--max_sec_tm', type=int, default=40, metavar='INT')
    parser.add_argument('--gc_clamp', type=int, default=1, metavar='INT')
    parser.add_argument('--max_end_gc', type=int, default=4, metavar='INT')
    parser.add_argument("--verbose", action="store_true",
                        help="Print runtime information to sys.stderr")
    return parser


def deduce(args, parser=None):
    """Fill conserved_left / conserved_right / diagnostic / amplicon (krisp_fasta.py:178-213); exit(1) if impossible."""
    def die():
        print("ERROR: Could not deduce input parameters", file=sys.stderr)
        if parser is not None:
            parser.print_help(sys.stderr)
        sys.exit(1)

    if args.amplicon is not None:
        if args.diagnostic is not None:
            args.conserved = (args.amplicon - args.diagnostic) // 2
            args.conserved_left = args.conserved
            args.conserved_right = args.conserved
        elif args.conserved is not None:
            args.diagnostic = args.amplicon - 2 * args.conserved
            args.conserved_left = args.conserved
            args.conserved_right = args.conserved
        elif (args.conserved_left is not None) and (args.conserved_right is not None):
            args.diagnostic = (args.amplicon - args.conserved_left - args.conserved_right)
        else:
            die()
    elif args.diagnostic is not None:
        if args.conserved is not None:
            args.amplicon = args.diagnostic + 2 * args.conserved
            args.conserved_left = args.conserved
            args.conserved_right = args.conserved
        elif (args.conserved_left is not None) and (args.conserved_right is not None):
            args.amplicon = args.diagnostic + args.conserved_left + args.conserved_right
        else:
            die()
    else:
        die()
    return args


def main(argv=None):
    parser = build_parser()
    args = deduce(parser.parse_args(sys.argv[1:] if argv is None else argv), parser)
    if args.primer3:
        try:
            import primer3  # noqa: F401
        except ImportError:
            print("ERROR: --primer3 needs the primer3-py package (the reference imports it unconditionally, Amplicon.py:3)", file=sys.stderr)
            sys.exit(1)
    # Primer3 settings travel as a dict (the reference parks them in a class attribute, krisp_fasta.py:218-221)
    p3_args = {k: v for k, v in vars(args).items()
               if k in ("tm", "gc", "primer_size", "amp_size", "max_sec_tm", "gc_clamp", "max_end_gc")}
    L, R = args.conserved_left, args.conserved_right
    D = args.amplicon - L - R           # the middle kstream's split [L, -R] really leaves (krisp_fasta.py:37)
    start_t = time.time()
    if args.verbose:
        print("Finding kmer-based diagnostic regions for:", file=sys.stderr)
        for i, filename in enumerate(args.files):
            print(f"({i}) {filename}", file=sys.stderr)
        print("With this as an outgroup:", file=sys.stderr)
        for i, filename in enumerate(args.outgroup):
            print(f"({i}) {filename}", file=sys.stderr)
        print(file=sys.stderr)
    res = search_files(args.files, args.outgroup, L, D, R, omit_soft=args.omit_soft,
                       want_records=args.out_align is not None)
    if args.verbose:
        print(f"Extracted {res.n_records:,} {args.amplicon}-kmers on the GPU", file=sys.stderr)
        print("Rendering output ... ", file=sys.stderr)
    ingroup = [simplename(f) for f in args.files] if len(args.outgroup) else None     # krisp_fasta.py:281-283
    found = render_output(res, getattr(res, "labels", None), ingroup=ingroup, out_csv=args.out_csv,
                          out_align=args.out_align, dot=args.dot_alignment, find_primers=bool(args.primer3), p3_args=p3_args)
    if args.verbose:
        print(f"=> Found {found:,} regions in {time.time() - start_t:.2f} s", file=sys.stderr)
    return 0


if __name__ == "__main__":
    sys.exit(main())
