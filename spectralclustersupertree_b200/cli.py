"""Command line front end, ``scs``.

Drop-in for the reference's command (ref: src/sc_supertree/cli.py:8-39): the options ``-i/--in-file``,
``-o/--out-file``, ``-p/--pcg-weighting`` (one | depth | branch | bootstrap, case-insensitive, default branch),
``--disable-contraction`` and ``--version`` behave the same, running without arguments prints the help, and
the supertree is written to the output file as Newick.  The work itself takes the flat route: the input file is
parsed natively into the source-tree store and handed to the native recursion, no node objects in between.
"""

from __future__ import annotations

import click

from . import __version__
from .load import load_forest
from .scs import WEIGHTINGS, supertree_of_forest

_CHOICES = ("one", "depth", "branch", "bootstrap")  # the order the reference lists them in
assert set(_CHOICES) == set(WEIGHTINGS)


def run(in_file: str, out_file: str, weighting: str = "branch", contract_edges: bool = True) -> None:
    """``load_trees`` + ``construct_supertree`` + ``write`` of the reference's command (ref: cli.py:33-39)."""
    forest = load_forest(in_file)
    if forest.num_trees == 0:  # what construct_supertree says about an empty list (ref: scs.py:63-65)
        msg = "There must be at least one tree to make a supertree."
        raise ValueError(msg)
    tree = supertree_of_forest(forest, weighting.lower(), contract_edges=contract_edges)
    tree.write(out_file)


def _main(in_file: str, out_file: str, pcg_weighting: str, disable_contraction: bool) -> None:
    run(in_file, out_file, pcg_weighting, contract_edges=not disable_contraction)


scs = click.Command(
    name="scs",
    callback=_main,
    no_args_is_help=True,
    help="Build the spectral cluster supertree of the source trees in a file.",
    params=[
        click.Option(["-i", "--in-file"], required=True, help="Line-separated Newick file with the source trees."),
        click.Option(["-o", "--out-file"], required=True, help="Where the supertree is written."),
        click.Option(
            ["-p", "--pcg-weighting"],
            type=click.Choice(list(_CHOICES), case_sensitive=False),
            default="branch",
            help="How edges of the proper cluster graph are weighted.",
        ),
        click.Option(["--disable-contraction"], is_flag=True, default=False,
                     help="Skip the edge contraction step (not recommended)."),
    ],
)
scs = click.version_option(__version__)(scs)


if __name__ == "__main__":
    scs()
