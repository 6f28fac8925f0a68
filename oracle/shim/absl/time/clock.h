// Empty stand-in: poller.cc includes absl/time/clock.h but calls nothing from it.
