def sync_global_devices(name=""):
    return None
