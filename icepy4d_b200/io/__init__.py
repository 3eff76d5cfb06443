from .export2bundler import write_bundler_out  # noqa: F401
