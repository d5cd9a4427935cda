from .main_cli import main_cli

main_cli()
