/*
 * ORACLE SHIM: token ids the lemon-generated grammar.h would define
 * (query/grammar.y:62-64 declares OR, AND, NOT first; the remaining
 * terminals are numbered in order of first appearance in the rules).
 * Only their distinctness matters to the shim parser and to the
 * token-stream checks modelled on tests/t_queryparser.c.
 */
#ifndef NXSB_ORACLE_SHIM_GRAMMAR_H
#define NXSB_ORACLE_SHIM_GRAMMAR_H

#define TOKEN_OR		1
#define TOKEN_AND		2
#define TOKEN_NOT		3
#define TOKEN_BR_OPEN		4
#define TOKEN_BR_CLOSE		5
#define TOKEN_FF_STRING		6
#define TOKEN_QUOTED_STRING	7

#endif
