def naturalsize(value, *args, **kwargs):
    return f"{value} B"
