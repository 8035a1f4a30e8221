__version__ = "2.6.1-stub"
