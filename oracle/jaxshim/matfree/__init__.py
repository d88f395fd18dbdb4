"""Stub of the third-party `matfree` package: only what `util/gp_util.py` imports at
module scope, so the reference's GP kernels can be imported.  `hutchinson.hutchinson`
is the estimator the reference's own `hutchinson._sample` restates
(`/root/reference/src/matfree_extensions/hutchinson.py:51-54`)."""
