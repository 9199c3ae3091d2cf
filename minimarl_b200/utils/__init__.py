from .sinks import MetricSink  # noqa: F401
