"""Drop-in replacements for the reference's keymask_ident modules (same module names, function
names, signatures, return values and on-disk artefacts); the arithmetic runs in the CUDA library.

Put this directory in front of the reference's on sys.path (or import it as a package) and keep
using the reference's own driver `main_keymask_ident.py` - or the equivalent one shipped here."""
