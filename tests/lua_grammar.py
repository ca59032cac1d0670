"""A Lua 5.1 grammar for lark (Earley), used to PARSE the LuaJIT glue in lua/radio_b200/ -- LuaJIT is not installed in the
build image, so this is as close to `luajit -bl` as the CPU test-suite gets.  The grammar follows the reference manual's
"complete syntax of Lua" (section 8 of the 5.1 manual) plus LuaJIT's `goto` / `::label::` (the reference uses them,
radio/core/composite.lua:194); operator precedence is spelled out in layers.  The one ambiguity of the language -- `f\n(g)(x)`:
a call continuation or a new statement that starts with a parenthesis -- is resolved the way Lua does (continuation) by giving
the parenthesised prefix expression a negative priority."""
from lark import Lark

LUA_GRAMMAR = r"""
start: block
block: stat* laststat?
?stat: ";"
     | varlist "=" explist                                   -> assign
     | functioncall
     | "do" block "end"                                      -> do_block
     | "while" exp "do" block "end"                          -> while_loop
     | "repeat" block "until" exp                            -> repeat_loop
     | "if" exp "then" block ("elseif" exp "then" block)* ("else" block)? "end"   -> if_stat
     | "for" NAME "=" exp "," exp ("," exp)? "do" block "end" -> for_num
     | "for" namelist "in" explist "do" block "end"          -> for_in
     | "function" funcname funcbody                          -> function_stat
     | "local" "function" NAME funcbody                      -> local_function
     | "local" namelist ("=" explist)?                       -> local_assign
     | "goto" NAME                                           -> goto_stat
     | "::" NAME "::"                                        -> label
laststat: "return" explist? ";"?   -> return_stat
        | "break" ";"?               -> break_stat
funcname: NAME ("." NAME)* method?
method: ":" NAME
varlist: var ("," var)*
namelist: NAME ("," NAME)*
explist: exp ("," exp)*

?exp: or_exp
?or_exp: and_exp ("or" and_exp)*
?and_exp: cmp_exp ("and" cmp_exp)*
?cmp_exp: cat_exp (CMP_OP cat_exp)*
?cat_exp: add_exp (".." cat_exp)?
?add_exp: mul_exp ((PLUS | MINUS) mul_exp)*
?mul_exp: unary_exp (MUL_OP unary_exp)*
?unary_exp: UNARY_OP unary_exp | MINUS unary_exp | pow_exp
?pow_exp: atom ("^" unary_exp)?
?atom: NIL | FALSE | TRUE | NUMBER | STRING | LONGSTRING | VARARG | function | prefixexp | tableconstructor

?prefixexp: var | functioncall | paren_exp
paren_exp.-10: "(" exp ")"
var: NAME | prefixexp "[" exp "]" | prefixexp "." NAME
functioncall: prefixexp args | prefixexp ":" NAME args
args: "(" explist? ")" | tableconstructor | STRING | LONGSTRING
function: "function" funcbody
funcbody: "(" parlist? ")" block "end"
parlist: namelist ("," VARARG)? | VARARG
tableconstructor: "{" fieldlist? "}"
fieldlist: field (("," | ";") field)* ("," | ";")?
field: "[" exp "]" "=" exp | NAME "=" exp | exp

CMP_OP: "<=" | ">=" | "==" | "~=" | "<" | ">"
NIL: "nil"
FALSE: "false"
TRUE: "true"
VARARG: "..."
PLUS: "+"
MINUS: "-"
MUL_OP: "*" | "/" | "%"
UNARY_OP: "not" | "#"
NAME: /(?!(?:and|break|do|else|elseif|end|false|for|function|goto|if|in|local|nil|not|or|repeat|return|then|true|until|while)\b)[A-Za-z_][A-Za-z_0-9]*/
NUMBER: /0[xX][0-9a-fA-F]+|(?:[0-9]+\.?[0-9]*|\.[0-9]+)(?:[eE][-+]?[0-9]+)?/
STRING: /"(?:\\.|[^"\\\n])*"|'(?:\\.|[^'\\\n])*'/
LONGSTRING: /\[\[.*?\]\]/s | /\[=\[.*?\]=\]/s
COMMENT: /--\[\[.*?\]\]/s | /--\[=\[.*?\]=\]/s | /--[^\n]*/
%import common.WS
%ignore WS
%ignore COMMENT
"""

_parser = None
_cache = {}


def parse_lua(text):
    """Parse a Lua chunk; raises lark.exceptions.LarkError on a syntax error."""
    global _parser
    if _parser is None:
        # the BASIC lexer (longest match): with the dynamic one a `-- comment` line can also be read as two unary minuses
        _parser = Lark(LUA_GRAMMAR, parser="earley", lexer="basic", ambiguity="resolve", propagate_positions=True)
    if text.startswith("#"):                      # shebang line
        text = "--" + text
    tree = _cache.get(text)                       # Earley is slow; the test-suite parses the same glue files many times
    if tree is None:
        tree = _cache[text] = _parser.parse(text)
    return tree


LUA_GLOBALS = {"require", "error", "tonumber", "tostring", "ipairs", "pairs", "os", "table", "string", "math", "setmetatable",
               "getmetatable", "type", "assert", "unpack", "print", "select", "next", "io", "rawget", "rawset", "pcall", "collectgarbage",
               "_G", "debug", "bit", "jit", "arg", "loadstring", "dofile", "package", "rawequal", "xpcall", "setfenv", "getfenv", "load"}


def undefined_globals(text):
    """Names READ or WRITTEN as variables that are neither local (declared earlier in an enclosing scope: `local`, function
    parameters, `self` of a method, loop variables) nor standard Lua / LuaJIT globals -- in Lua a misspelled local silently
    becomes a nil global, the classic way for never-executed glue to be wrong.  Returns a sorted list of (name, line)."""
    from lark import Token, Tree
    tree = parse_lua(text)
    scopes = [set()]
    bad = set()

    def declared(name):
        return any(name in s for s in scopes) or name in LUA_GLOBALS

    def names_of(node):                      # namelist -> [names]
        return [str(t) for t in node.children if isinstance(t, Token) and t.type == "NAME"]

    def visit(node):
        if isinstance(node, Token):
            return
        d = node.data
        if d == "block":
            scopes.append(set())
            for c in node.children:
                visit(c)
            scopes.pop()
        elif d == "local_assign":
            for c in node.children[1:]:
                visit(c)                     # the initialisers see the OLD scope
            scopes[-1].update(names_of(node.children[0]))
        elif d == "local_function":
            scopes[-1].add(str(node.children[0]))
            visit_funcbody(node.children[1], False)
        elif d == "function_stat":
            fn = node.children[0]
            toks = [t for t in fn.children if isinstance(t, Token)]
            if not declared(str(toks[0])):
                bad.add((str(toks[0]), toks[0].line))
            method = any(isinstance(c, Tree) and c.data == "method" for c in fn.children)
            visit_funcbody(node.children[1], method)
        elif d == "function":
            visit_funcbody(node.children[0], False)
        elif d == "for_num":
            exps, body = node.children[1:-1], node.children[-1]
            for e in exps:
                visit(e)
            scopes.append({str(node.children[0])})
            visit(body)
            scopes.pop()
        elif d == "for_in":
            visit(node.children[1])
            scopes.append(set(names_of(node.children[0])))
            visit(node.children[2])
            scopes.pop()
        elif d == "repeat_loop":             # the `until` expression sees the block's locals
            scopes.append(set())
            blk, cond = node.children
            for c in blk.children:
                visit(c)
            visit(cond)
            scopes.pop()
        elif d == "var":
            if len(node.children) == 1 and isinstance(node.children[0], Token):
                t = node.children[0]
                if not declared(str(t)):
                    bad.add((str(t), t.line))
            else:
                visit(node.children[0])
                if isinstance(node.children[1], Tree):
                    visit(node.children[1])  # index expression; a `.NAME` field is not a variable
        elif d == "functioncall":
            for c in node.children:
                if isinstance(c, Tree):
                    visit(c)                 # the method NAME of a `:` call is a token, skipped
        elif d == "field":
            kids = node.children
            if len(kids) == 2 and isinstance(kids[0], Token) and kids[0].type == "NAME":
                visit(kids[1])               # NAME = exp: the key is not a variable
            else:
                for c in kids:
                    visit(c)
        elif d in ("goto_stat", "label", "funcname", "namelist"):
            return
        else:
            for c in node.children:
                visit(c)

    def visit_funcbody(fb, method):
        params = set(["self"]) if method else set()
        block = fb.children[-1]
        for c in fb.children[:-1]:
            if isinstance(c, Tree) and c.data == "parlist":
                for cc in c.children:
                    if isinstance(cc, Tree) and cc.data == "namelist":
                        params.update(names_of(cc))
        scopes.append(params)
        visit(block)
        scopes.pop()

    visit(tree)
    return sorted(bad)
