"""Drop-in `runners` package: same module and class names as the reference's runners/ directory."""
