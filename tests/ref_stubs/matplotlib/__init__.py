"""Import-only stand-in (test infrastructure): train.py imports matplotlib.pyplot / cm for debug plots it never draws here."""
