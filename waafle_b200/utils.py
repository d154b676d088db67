"""Generic front-end helpers with the reference's user-visible behaviour.

Same messages and exit behaviour as reference waafle/utils.py:46-143 (say / die /
try_open / read_contig_lengths / write_rowdict); written fresh for this package.
"""

import bz2
import gzip
import io
import sys
from collections import OrderedDict


def say(*args):
    print(" ".join(str(a) for a in args), file=sys.stderr)


def die(*args):
    """Reference error contract (UT:49-52): message on stderr, then sys.exit("EXITING.")."""
    say("LETHAL ERROR:", *args)
    sys.exit("EXITING.")


def try_open(path, mode="r"):
    """Open plain / .gz / .bz2 files as text; exit with the reference's message on failure."""
    try:
        if path.endswith(".gz"):
            return io.TextIOWrapper(gzip.open(path, mode + "b"))
        if path.endswith(".bz2"):
            return io.TextIOWrapper(bz2.open(path, mode + "b"))
        return open(path, mode)
    except Exception:
        sys.exit("Can't open file: {}".format(path))


def read_contig_lengths(fasta):
    """Contig name -> length in file order (UT:109-120)."""
    data = OrderedDict()
    header = None
    with try_open(fasta) as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            if line[0] == ">":
                header = line[1:].split()[0]
                data[header] = 0
            else:
                data[header] += len(line)
    return data


def format_field(value, precision=4, empty_field="--"):
    """One TSV cell exactly as UT:137-139 prints it."""
    if isinstance(value, float):
        value = "{:.{p}f}".format(value, p=precision)
    value = str(value)
    return value if value != "" else empty_field


def write_row(values, file, delim="\t"):
    print(delim.join(values), file=file)
