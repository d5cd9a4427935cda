"""Command-line front end for the in-scope subcommands (same names and flags as `finaletoolkit`)."""
from .main_cli import main_cli

__all__ = ["main_cli"]
