/* empty GSL stand-in: the hot-path sources only need the include to resolve */
