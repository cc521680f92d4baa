from .panorama_to_plane_pitch import cli

cli()
