"""Command line front end, ``scs``.

Drop-in for the reference's command (ref: src/sc_supertree/cli.py:8-39): same options, defaults, help strings and
behaviour (``no_args_is_help``, ``--version``, case-insensitive weighting choice); the supertree is written to the
output file as Newick.  The work itself takes the flat route: the input file is parsed natively into the
source-tree store, the native recursion runs on the GPU, and the result is written as Newick natively -- no node
object is built at either end (``run``).
"""

from __future__ import annotations

from pathlib import Path
from typing import Literal

import click

from . import __version__
from .load import load_forest
from .scs import supertree_newick_of_forest


def run(in_file: str, out_file: str, weighting: str = "branch", contract_edges: bool = True) -> None:
    """``load_trees`` + ``construct_supertree`` + ``write`` of the reference's command (ref: cli.py:33-39)."""
    forest = load_forest(in_file)
    if forest.num_trees == 0:  # what construct_supertree says about an empty list (ref: scs.py:63-65)
        msg = "There must be at least one tree to make a supertree."
        raise ValueError(msg)
    text = supertree_newick_of_forest(forest, weighting.lower(), contract_edges=contract_edges)
    Path(out_file).write_text(text + "\n")


@click.command(no_args_is_help=True)
@click.version_option(__version__)
@click.option("-i", "--in-file", required=True, help="File containing source trees.")
@click.option("-o", "--out-file", required=True, help="Output file.")
@click.option(
    "-p",
    "--pcg-weighting",
    help="Proper cluster graph weighting strategy.",
    default="branch",
    type=click.Choice(["one", "depth", "branch", "bootstrap"], case_sensitive=False),
)
@click.option(
    "--disable-contraction",
    help="Disable edge contraction (not recommended).",
    default=False,
    is_flag=True,
)
def scs(
    in_file: str,
    out_file: str,
    pcg_weighting: Literal["one", "depth", "branch", "bootstrap"],
    *,
    disable_contraction: bool,
) -> None:
    """Run spectral cluster supertree over the given set of source trees."""
    run(in_file, out_file, pcg_weighting, contract_edges=not disable_contraction)


if __name__ == "__main__":
    scs()
