"""The ``scs`` command (ref: src/sc_supertree/cli.py:8-39): same options, defaults and output."""

from __future__ import annotations

from typing import Literal

import click

from . import __version__
from .load import load_forest
from .scs import supertree_of_forest


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
    # load_trees + construct_supertree (ref: cli.py:33-38) without the detour over node objects
    forest = load_forest(in_file)
    if forest.num_trees == 0:  # ref: scs.py:63-65
        msg = "There must be at least one tree to make a supertree."
        raise ValueError(msg)
    supertree = supertree_of_forest(
        forest,
        pcg_weighting.lower(),
        contract_edges=not disable_contraction,
    )
    supertree.write(out_file)


if __name__ == "__main__":
    scs()
