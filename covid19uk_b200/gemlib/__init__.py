"""Stand-ins for the gemlib symbols the reference imports, backed by the
CUDA library: ``gemlib.distributions``, ``gemlib.util``, ``gemlib.mcmc``."""
