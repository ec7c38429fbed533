"""Command-line drop-ins for the reference's run_inference_custom.py entry points (CLI row)."""
