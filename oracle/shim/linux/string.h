/* oracle/shim -- empty stand-in for the kernel header the reference includes */
