"""B200 drop-ins for the numeric pieces of gccNMF/realtime (the GUI, PyAudio I/O and process orchestration of the
reference are out of scope: SURVEY.md section 2)."""
