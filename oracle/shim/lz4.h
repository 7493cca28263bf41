/* oracle/shim -- liblz4 block API is not used by the compiled reference path */
