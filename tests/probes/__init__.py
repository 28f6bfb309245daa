"""Diagnostic scripts that use the CPU oracle (test infrastructure: only tests/ may touch oracle/). Not collected by pytest."""
