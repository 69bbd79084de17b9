"""Plugin loader, ABI driver, NIC discovery, telemetry readers and env helpers."""
from .native import (base64, chunk_count, chunk_size, config, exec_stats, find_interfaces, if_filter_accepts,  # noqa: F401
                     load, metrics_text, parse_user_pass_and_addr, reload_config, sockaddr_roundtrip,
                     telemetry_flush, trace_json, version)
