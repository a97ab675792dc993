"""Post-GWAS joins that close the two-pass loop: which gene clusters carry associated patterns
(`panfeed-get-clusters`, /root/reference/panfeed/get_clusters.py:71-101) and the positional
annotation of their k-mers (`panfeed-get-kmers`, get_kmers.py:88-145).  Same options, same output.

What the reference spends its time on — streaming kmers_to_hashes.tsv / kmers.tsv through pandas
100,000 rows at a time and keeping the rows whose hash / cluster is in a set — is done by the
library's host threads (`pf_tsv_filter`); the small joins and the output formatting that follow
are pandas, as in the reference, so the text is the same.  Compressed (.gz) inputs cannot be
mapped and take the pandas route.
"""
import argparse
import io
import logging
import sys

import pandas as pd

from . import __version__, capi

logger = logging.getLogger("panfeed")


def set_logging(v):
    logger.propagate = True
    logger.setLevel(logging.DEBUG)
    ch = logging.StreamHandler()
    ch.setLevel(logging.INFO if v == 0 else logging.DEBUG)
    ch.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(message)s", "%H:%M:%S"))
    logger.addHandler(ch)


def _common(parser):
    parser.add_argument("-a", "--associations", required=True,
                        help="TSV file containing hashes and their significance (e.g. pyseer output; "
                             "tab-delimited, with a header, first column the hash, another column - by "
                             "default 'lrt-pvalue' - the association p-value)")
    parser.add_argument("-p", "--kmers-to-hashes", required=True,
                        help="TSV file relating gene clusters, kmers, and their hashes "
                             "(i.e. panfeed's kmers_to_hashes.tsv file)")


def _tail(parser):
    parser.add_argument("-v", action="count", default=0, help="Increase verbosity level")
    parser.add_argument("--version", action="version", version="%(prog)s " + __version__)


def passing_associations(path, column, threshold, output=None):
    """get_clusters.py:76-88 / get_kmers.py:93-106: the associations at or under the threshold."""
    a = pd.read_csv(path, sep="\t", index_col=0)
    if column not in a.columns:
        logger.warning(f"Associations file does not have the {column} column")
        sys.exit(1)
    a = a[a[column] <= threshold]
    if output is not None:
        a.to_csv(output, sep="\t")
        logger.info(f"Saved filtered associations to {output}")
    return a


def filter_rows(path, column_name, keys, n_threads=0):
    """The rows of a TSV whose `column_name` is in `keys`, as a DataFrame with the file's header:
    pd.concat([x[x[column_name].isin(keys)] for x in chunks]) of the reference, scanned natively."""
    if str(path).endswith(".gz"):
        chunks = pd.read_csv(path, sep="\t", iterator=True, chunksize=100_000)
        return pd.concat([x[x[column_name].isin(keys)] for x in chunks])
    with open(path, "rb") as fh:
        header = fh.readline()
    names = header.decode().rstrip("\r\n").split("\t")
    if column_name not in names:
        raise KeyError(column_name)
    body, _ = capi.tsv_filter(path, names.index(column_name), [str(k) for k in keys], True, n_threads)
    return pd.read_csv(io.BytesIO(header + body), sep="\t")


def get_clusters_main(argv=None):
    parser = argparse.ArgumentParser(description="Indicate which genes clusters have significantly "
                                                 "associated patterns")
    _common(parser)
    parser.add_argument("-t", "--threshold", type=float, default=1,
                        help="Association p-value threshold (default %(default).2f)")
    parser.add_argument("-c", "--column", default="lrt-pvalue",
                        help="P-value column in the associations file (default %(default)s)")
    parser.add_argument("-o", "--output", default=None,
                        help="Filename to save filtered associations table (not saved by default)")
    _tail(parser)
    args = parser.parse_args(argv)
    set_logging(args.v)
    a = passing_associations(args.associations, args.column, args.threshold, args.output)
    passing = set(a.index)
    logger.info(f"{len(passing)} patterns pass the association threshold")
    h = filter_rows(args.kmers_to_hashes, "hashed_pattern", passing)
    clusters = set(h["cluster"].unique())
    logger.info(f"Found significant associations for {len(clusters)} gene clusters")
    for c in clusters:
        print(c)


def get_kmers_main(argv=None):
    parser = argparse.ArgumentParser(description="Annotate association results with positional information")
    _common(parser)
    parser.add_argument("-k", "--kmers", required=True,
                        help="TSV file with positional information of individual k-mers "
                             "(i.e. panfeed's kmers.tsv file)")
    parser.add_argument("-t", "--threshold", type=float, default=1,
                        help="Association p-value threshold (default %(default).2f)")
    parser.add_argument("-c", "--column", default="lrt-pvalue",
                        help="P-value column in the associations file (default %(default)s)")
    parser.add_argument("-o", "--output", default=None,
                        help="Filename to save filtered associations table (not saved by default)")
    parser.add_argument("--only-passing", action="store_true", default=False,
                        help="Only output passing k-mers (default is all)")
    parser.add_argument("--clusters-per-iteration", type=int, default=15,
                        help="Number of clusters to be considered in each iteration, a higher number means "
                             "faster execution but higher memory usage (default %(default)d)")
    _tail(parser)
    args = parser.parse_args(argv)
    set_logging(args.v)
    a = passing_associations(args.associations, args.column, args.threshold, args.output)
    a.index.name = "hashed_pattern"
    passing = set(a.index)
    logger.info(f"{len(passing)} patterns pass the p-value threshold {args.threshold}")
    h = filter_rows(args.kmers_to_hashes, "hashed_pattern", passing).set_index("hashed_pattern")
    clusters = list(set(h["cluster"].unique()))
    kmers = set(h["k-mer"].unique())
    logger.info(f"Found {len(clusters)} gene clusters")
    logger.info(f"Found {len(kmers)} k-mers")
    first = True
    # a limited number of clusters per pass over kmers.tsv, as in the reference (memory)
    for i, at in enumerate(range(0, len(clusters), args.clusters_per_iteration)):
        bunch = clusters[at:at + args.clusters_per_iteration]
        logger.info(f"Searching for k-mers for {len(bunch)} clusters (iteration {i + 1})")
        k = filter_rows(args.kmers, "cluster", bunch).set_index(["cluster", "k-mer"])
        b = a.join(h, how="inner")
        b = b.reset_index().set_index(["cluster", "k-mer"]).join(k, how="left" if args.only_passing else "right")
        b.to_csv(sys.stdout, sep="\t", header=first)
        first = False


if __name__ == "__main__":
    get_clusters_main()
