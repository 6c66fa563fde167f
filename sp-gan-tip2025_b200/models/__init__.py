"""Drop-in mirrors of the reference's hot-path modules (same names, constructor / forward signatures and
state-dict keys): `models.custom_ops`, `models.spherenet`, `models.ops`, `models.spgan_ops`, `models.spgan_ops_gs`.
A reference checkout uses them by putting this directory's parent first on sys.path."""
