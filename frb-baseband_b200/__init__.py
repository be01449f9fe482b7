"""b2f: B200-native baseband(VDIF) -> SIGPROC filterbank, drop-in for the digifil+splice
stage of pharaofranz/frb-baseband (process_vdif.py / base2fil.sh:395-448)."""
__version__ = "0.1.0"
