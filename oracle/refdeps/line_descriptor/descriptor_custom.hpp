// Stand-in: the hot path does not touch the line-descriptor module that plsvo/global.h pulls in.
#ifndef PLSVO_REFDEPS_LINE_DESCRIPTOR
#define PLSVO_REFDEPS_LINE_DESCRIPTOR
#endif
