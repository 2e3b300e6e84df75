"""IVF / IVFPQ index families (placeholder until the inverted-list kernels land)."""


def create(owner, kind):
    raise NotImplementedError(f"index_type={kind!r}: the IVF kernels are not built yet in this revision; "
                              "use index_type='Flat'")
